/* oracle/shims/boost/stacktrace.hpp -- empty stand-in (TEST INFRASTRUCTURE ONLY). */
#ifndef NTS_ORACLE_SHIM_BOOST_STACKTRACE
#define NTS_ORACLE_SHIM_BOOST_STACKTRACE
#include <ostream>
namespace boost { namespace stacktrace { struct stacktrace {}; inline std::ostream &operator<<(std::ostream &o, const stacktrace &) { return o; } } }
#endif
