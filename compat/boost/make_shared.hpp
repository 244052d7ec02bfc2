// boost/shared_ptr.hpp stand-in (Boost is not installed in this image): boost::shared_ptr is std::shared_ptr.
#pragma once
#include <memory>
namespace boost {
template <class T> using shared_ptr = std::shared_ptr<T>;
template <class T, class... A> shared_ptr<T> make_shared(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }
template <class T, class U> shared_ptr<T> dynamic_pointer_cast(const shared_ptr<U>& p) { return std::dynamic_pointer_cast<T>(p); }
template <class T, class U> shared_ptr<T> static_pointer_cast(const shared_ptr<U>& p) { return std::static_pointer_cast<T>(p); }
}  // namespace boost
