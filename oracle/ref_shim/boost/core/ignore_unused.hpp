// Minimal stand-in for <boost/core/ignore_unused.hpp> (oracle build only).
#pragma once
namespace boost {
template <class... Ts> inline void ignore_unused(Ts const &...) {}
}
