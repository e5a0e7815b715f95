/* Minimal stand-in for <boost/timer/timer.hpp> (TEST INFRASTRUCTURE ONLY):
 * a wall-clock cpu_timer on std::chrono::steady_clock. */
#ifndef CSM_ORACLE_BOOST_TIMER_SHIM
#define CSM_ORACLE_BOOST_TIMER_SHIM

#include <chrono>
#include <cstdint>

namespace boost {
namespace timer {

using nanosecond_type = std::int_least64_t;

struct cpu_times
{
    nanosecond_type wall;
    nanosecond_type user;
    nanosecond_type system;
};

class cpu_timer
{
public:
    cpu_timer() { this->start(); }

    bool is_stopped() const { return this->mStopped; }

    cpu_times elapsed() const
    {
        nanosecond_type total = this->mAccumulated;
        if (!this->mStopped)
            total += Since(this->mStart);
        return cpu_times { total, 0, 0 };
    }

    void start()
    {
        this->mAccumulated = 0;
        this->mStopped = false;
        this->mStart = Clock::now();
    }

    void stop()
    {
        if (this->mStopped)
            return;
        this->mAccumulated += Since(this->mStart);
        this->mStopped = true;
    }

    void resume()
    {
        if (!this->mStopped)
            return;
        this->mStopped = false;
        this->mStart = Clock::now();
    }

private:
    using Clock = std::chrono::steady_clock;

    static nanosecond_type Since(const Clock::time_point& start)
    {
        return std::chrono::duration_cast<std::chrono::nanoseconds>(
            Clock::now() - start).count();
    }

    Clock::time_point mStart;
    nanosecond_type   mAccumulated = 0;
    bool              mStopped = false;
};

} /* namespace timer */
} /* namespace boost */

#endif /* CSM_ORACLE_BOOST_TIMER_SHIM */
