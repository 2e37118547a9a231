// pairhmm_tables.h -- host-built constant tables of the PairHMM engine.
//
// Same construction as the reference's Context<float>/Context<double>
// (/root/reference/src/main/native/pairhmm/Context.h:65-89 Jacobian + matchToMatch tables,
// :133-148 and :174-189 ph2pr / INITIAL_CONSTANT), evaluated with the host libm so that the
// constants are bit-identical to the ones GKL computes in the same process.  Only the entries
// reachable with quals & 127 are kept (128 ph2pr values, 128*129/2 matchToMatch values).
#pragma once

namespace gklb {

constexpr int kPh2prSize = 128;
constexpr int kMmSize = (128 * 129) / 2;

struct HostTables {
  float ph2pr_f[kPh2prSize];
  float mm_f[kMmSize];
  double ph2pr_d[kPh2prSize];
  double mm_d[kMmSize];
  float init_f;         // 2^120
  float log10_init_f;   // log10f(2^120)
  double init_d;        // 2^1020
  double log10_init_d;  // log10(2^1020)
};

const HostTables& host_tables();  // built once, thread-safe

}  // namespace gklb
