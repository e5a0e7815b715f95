/* Minimal stand-in for <boost/property_tree/ptree.hpp> (TEST INFRASTRUCTURE ONLY).
 * Boost is not installed in this image; the reference metric registry only
 * builds trees that nobody reads in the oracle. All operations are no-ops. */
#ifndef CSM_ORACLE_BOOST_PTREE_SHIM
#define CSM_ORACLE_BOOST_PTREE_SHIM

#include <string>
#include <utility>

namespace boost {
namespace property_tree {

struct ptree
{
    template <typename T>
    ptree& put(const std::string&, const T&) { return *this; }
    template <typename T>
    ptree& put_value(const T&) { return *this; }
    ptree& add_child(const std::string&, const ptree&) { return *this; }
    ptree& put_child(const std::string&, const ptree&) { return *this; }
    void push_back(const std::pair<std::string, ptree>&) { }
    template <typename T>
    T get(const std::string&) const { return T(); }
    template <typename T>
    T get(const std::string&, const T& defaultValue) const
    { return defaultValue; }
    const ptree& get_child(const std::string&) const { return *this; }
};

} /* namespace property_tree */
} /* namespace boost */

#endif /* CSM_ORACLE_BOOST_PTREE_SHIM */
