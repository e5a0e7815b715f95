/* metrics.hpp -- where the adapters report the reference's metric ids.
 *
 * The reference registers value sequences "<name>.<Metric>" with its MetricManager singleton
 * (metric/metric.hpp, e.g. scan_matcher_correlative.cpp:38-70, loop_detector_branch_bound.cpp:24-38)
 * and downstream tools read them from the saved *.metric.json. The adapters observe the same ids
 * with the same meaning through this interface; an integration forwards Observe() to
 * MetricManager::Instance()->AddValueSequence<T>(id)->Observe(v), a test records
 * them. No process-global state: every matcher / detector holds its own sink (null = off). */
#pragma once

#include <chrono>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace csm_host {

class MetricSink
{
public:
    virtual ~MetricSink() = default;
    virtual void Observe(const std::string& id, double value) = 0;
};
using MetricSinkPtr = std::shared_ptr<MetricSink>;

/* Keeps every observed value in order, per id (what a ValueSequence does) */
class MetricRecorder final : public MetricSink
{
public:
    void Observe(const std::string& id, double value) override { mValues[id].push_back(value); }
    const std::map<std::string, std::vector<double>>& Values() const { return mValues; }
    void Clear() { mValues.clear(); }

private:
    std::map<std::string, std::vector<double>> mValues;
};

/* Metric::Timer (metric/metric.hpp): wall clock in microseconds */
class MicroTimer
{
public:
    MicroTimer() { Start(); }
    void Start() { mStart = std::chrono::steady_clock::now(); }
    double ElapsedMicro() const
    {
        return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - mStart).count();
    }

private:
    std::chrono::steady_clock::time_point mStart;
};

} /* namespace csm_host */
