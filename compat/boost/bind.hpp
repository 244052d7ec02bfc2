// boost/bind.hpp stand-in: boost::bind -> std::bind, _1.. -> std::placeholders.
#pragma once
#include <functional>
namespace boost {
template <class... A> auto bind(A&&... a) -> decltype(std::bind(std::forward<A>(a)...)) { return std::bind(std::forward<A>(a)...); }
}  // namespace boost
using std::placeholders::_1;
using std::placeholders::_2;
using std::placeholders::_3;
