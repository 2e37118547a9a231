// pairhmm_tables.cc -- see pairhmm_tables.h.
#include "pairhmm_tables.h"

#include <algorithm>
#include <cmath>
#include <vector>

namespace gklb {
namespace {

constexpr int kMaxQual = 127;               // quals are masked with 127 before every lookup
constexpr double kJacTolerance = 8.0;       // MAX_JACOBIAN_TOLERANCE
constexpr double kJacStep = 0.0001;         // JACOBIAN_LOG_TABLE_STEP
constexpr double kJacInvStep = 1.0 / kJacStep;
constexpr int kJacSize = (int)(kJacTolerance / kJacStep) + 1;

// log10(10^a + 10^b) through the 1e-4-step Jacobian table, hard rounding of the index
// (Context.h:91-122).  T is the precision the reference instantiates the context with.
template <class T>
T approx_log10_sum(T small, T big, const std::vector<T>& jac) {
  if (small > big) std::swap(small, big);
  if (std::isinf(small) || std::isinf(big)) return big;
  const T diff = big - small;
  if (diff >= (T)kJacTolerance) return big;
  const T v = (T)(diff * ((T)kJacInvStep));
  const int ind = (v > (T)0.0) ? (int)(v + (T)0.5) : (int)(v - (T)0.5);
  return big + jac[ind];
}

template <class T>
void build_mm(T* mm) {
  std::vector<T> jac(kJacSize);
  for (int k = 0; k < kJacSize; k++) jac[k] = (T)(std::log10(1.0 + std::pow(10.0, -((double)k) * kJacStep)));
  const double kInvLn10 = 0.434294;  // the reference's truncated constant (Context.h:78)
  for (int i = 0, offset = 0; i <= kMaxQual; offset += ++i) {
    for (int j = 0; j <= i; j++) {
      const double log10_sum = approx_log10_sum<T>((T)-0.1 * (T)i, (T)-0.1 * (T)j, jac);
      const double m2m_log10 = std::log1p(-std::min(1.0, std::pow(10, log10_sum))) * kInvLn10;
      mm[offset + j] = (T)(std::pow(10, m2m_log10));
    }
  }
}

HostTables* make_tables() {
  HostTables* t = new HostTables;
  build_mm<float>(t->mm_f);
  build_mm<double>(t->mm_d);
  for (int x = 0; x < kPh2prSize; x++) {
    t->ph2pr_d[x] = std::pow(10.0, -((double)x) / 10.0);
    t->ph2pr_f[x] = powf(10.f, -((float)x) / 10.f);
  }
  t->init_d = std::ldexp(1.0, 1020);
  t->log10_init_d = std::log10(t->init_d);
  t->init_f = ldexpf(1.f, 120);
  t->log10_init_f = log10f(t->init_f);
  return t;
}

}  // namespace

const HostTables& host_tables() {
  static const HostTables* t = make_tables();
  return *t;
}

}  // namespace gklb
