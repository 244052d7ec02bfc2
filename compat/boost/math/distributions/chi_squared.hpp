// boost/math/distributions/chi_squared.hpp stand-in: quantile(chi_squared(dof), p) as gtsam/chi2.h:17-26 uses it
// (regularised lower incomplete gamma by series / continued fraction, inverted by bisection; 1e-12 relative).
#pragma once
#include <cmath>
namespace boost { namespace math {
class chi_squared {
  double k_;
 public:
  explicit chi_squared(double dof) : k_(dof) {}
  double degrees_of_freedom() const { return k_; }
};
namespace detail_chi2 {
inline double gamma_p(double a, double x) {        // P(a, x)
  if (x <= 0) return 0.0;
  const double lg = std::lgamma(a);
  if (x < a + 1.0) {
    double sum = 1.0 / a, term = sum;
    for (int n = 1; n < 10000; ++n) { term *= x / (a + n); sum += term; if (std::fabs(term) < std::fabs(sum) * 1e-16) break; }
    return sum * std::exp(-x + a * std::log(x) - lg);
  }
  double b = x + 1.0 - a, c = 1e300, d = 1.0 / b, h = d;
  for (int i = 1; i < 10000; ++i) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b; if (std::fabs(d) < 1e-300) d = 1e-300;
    c = b + an / c; if (std::fabs(c) < 1e-300) c = 1e-300;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < 1e-16) break;
  }
  return 1.0 - std::exp(-x + a * std::log(x) - lg) * h;
}
}  // namespace detail_chi2
inline double cdf(const chi_squared& d, double x) { return detail_chi2::gamma_p(0.5 * d.degrees_of_freedom(), 0.5 * x); }
inline double quantile(const chi_squared& d, double p) {
  double lo = 0.0, hi = d.degrees_of_freedom() + 10.0;
  while (cdf(d, hi) < p) hi *= 2.0;
  for (int i = 0; i < 200; ++i) { const double mid = 0.5 * (lo + hi); if (cdf(d, mid) < p) lo = mid; else hi = mid; }
  return 0.5 * (lo + hi);
}
}}  // namespace boost::math
