// Arithmetic of the GPU-resident genetic operators: simulated binary crossover, polynomial
// mutation (real and integer variants: operators.py:66-78), binary tournament, duplicate elimination and the
// NSGA-II / GA survival (run.py:59-68 get_algorithm("nsga2" | "ga")).  The reference takes all of it from pymoo
// 0.4.2.1; the arithmetic here follows clip_glass_b200/ga.py operation by operation (same IEEE operations in the
// same order, multiplications and additions kept un-fused), so that with the same uniform draws the children differ
// from the host operators only by the last bit of pow().
//
// Everything is written as __host__ __device__ element / phase functions: ga.cu wraps them in kernels, and
// tests/native/ga_host.cpp compiles the SAME functions with g++ so that the CPU suite checks the logic against
// ga.py without a GPU (test infrastructure only: nothing in the product loads that harness).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GA_HD __host__ __device__ __forceinline__
#else
#define GA_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define GA_MUL(a, b) __dmul_rn((a), (b))
#define GA_ADD(a, b) __dadd_rn((a), (b))
#define GA_SUB(a, b) __dsub_rn((a), (b))
#define GA_DIV(a, b) __ddiv_rn((a), (b))
#define GA_FOR(j, n) for (int j = (int)threadIdx.x; j < (n); j += (int)blockDim.x)
#define GA_SYNC() __syncthreads()
#define GA_TID ((int)threadIdx.x)
#define GA_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define GA_MULHI(a, b) __umulhi((a), (b))
#else
#define GA_MUL(a, b) ((a) * (b))
#define GA_ADD(a, b) ((a) + (b))
#define GA_SUB(a, b) ((a) - (b))
#define GA_DIV(a, b) ((a) / (b))
#define GA_FOR(j, n) for (int j = 0; j < (n); ++j)
#define GA_SYNC() ((void)0)
#define GA_TID 0
#define GA_ATOMIC_ADD(p, v) (*(p) += (v))
#define GA_MULHI(a, b) ((uint32_t)(((uint64_t)(a) * (uint64_t)(b)) >> 32))
#endif

namespace glass_ga {

// ---------------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter = (lo, hi, stream, 0), key = seed.  Two 53-bit uniforms in [0, 1) per
// counter value.
// ---------------------------------------------------------------------------------------------------------------
GA_HD void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = GA_MULHI(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = GA_MULHI(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
constexpr uint32_t kStreamTag = 0x47414F50u;   // "GAOP": keeps these draws apart from the noise generator's
GA_HD double u53(uint32_t a, uint32_t b) {
  return (double)((((uint64_t)(a >> 5)) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}
// uniforms number 2q and 2q+1 of the stream (seed, offset)
GA_HD void uniform_pair(uint64_t seed, uint64_t q, double& u0, double& u1) {
  uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32), kStreamTag, 0u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  u0 = u53(c[0], c[1]);
  u1 = u53(c[2], c[3]);
}

// ---------------------------------------------------------------------------------------------------------------
// operators (ga.py SimulatedBinaryCrossover._do / .do, PolynomialMutation._do, _IntegerFromFloat.do)
// ---------------------------------------------------------------------------------------------------------------
struct OpParams {
  double sbx_eta, sbx_prob, sbx_prob_var;   // operators.py:69 real_sbx(prob=1.0, eta=3.0); prob_per_variable 0.5
  double pm_eta, pm_prob;                   // operators.py:70 real_pm(prob=0.5, eta=3.0); pm_prob < 0 => 1 / n_var
  int32_t n_var;
  int32_t integer;                          // operators.py:75-77 int_sbx / int_pm: round + clip after each operator
};

GA_HD double clipd(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

GA_HD double sbx_betaq(double beta, double rand, double eta) {
  const double alpha = GA_SUB(2.0, pow(beta, -GA_ADD(eta, 1.0)));
  const double e = GA_DIV(1.0, GA_ADD(eta, 1.0));
  if (rand <= GA_DIV(1.0, alpha)) return pow(GA_MUL(rand, alpha), e);
  return pow(GA_DIV(1.0, GA_SUB(2.0, GA_MUL(rand, alpha))), e);
}

// One variable of one mating: parents x0 / x1 -> children c0 / c1.  r_do / r_u / r_swap: this variable's draws,
// r_keep: the mating's draw (a mating is recombined with probability sbx_prob).
GA_HD void sbx_element(const OpParams& p, double x0, double x1, double xl, double xu, double r_do, double r_u,
                       double r_swap, double r_keep, double& c0, double& c1) {
  if (r_keep >= p.sbx_prob) {            // off[:, keep] = X[:, keep]: the parents, not clipped
    c0 = x0;
    c1 = x1;
    return;
  }
  const bool rec = (r_do <= p.sbx_prob_var) && (fabs(GA_SUB(x0, x1)) > 1e-14);
  if (rec) {
    const double y1 = fmin(x0, x1), y2 = fmax(x0, x1);
    const double delta = fmax(GA_SUB(y2, y1), 1e-10);
    const double sum = GA_ADD(y1, y2);
    const double b1 = GA_ADD(1.0, GA_DIV(GA_MUL(2.0, GA_SUB(y1, xl)), delta));
    const double b2 = GA_ADD(1.0, GA_DIV(GA_MUL(2.0, GA_SUB(xu, y2)), delta));
    const double a = GA_MUL(0.5, GA_SUB(sum, GA_MUL(sbx_betaq(b1, r_u, p.sbx_eta), delta)));
    const double b = GA_MUL(0.5, GA_ADD(sum, GA_MUL(sbx_betaq(b2, r_u, p.sbx_eta), delta)));
    const bool swap = r_swap <= 0.5;
    c0 = swap ? b : a;
    c1 = swap ? a : b;
  } else {
    c0 = x0;
    c1 = x1;
  }
  c0 = clipd(c0, xl, xu);
  c1 = clipd(c1, xl, xu);
}

GA_HD double pm_element(const OpParams& p, double x, double xl, double xu, double r_do, double r_u) {
  const double prob = p.pm_prob >= 0.0 ? p.pm_prob : GA_DIV(1.0, (double)p.n_var);
  if (!(r_do < prob)) return x;
  const double span = GA_SUB(xu, xl);
  const double d1 = GA_DIV(GA_SUB(x, xl), span), d2 = GA_DIV(GA_SUB(xu, x), span);
  const double e1 = GA_ADD(p.pm_eta, 1.0), mp = GA_DIV(1.0, e1);
  double dq;
  if (r_u <= 0.5) {
    const double v = GA_ADD(GA_MUL(2.0, r_u), GA_MUL(GA_SUB(1.0, GA_MUL(2.0, r_u)), pow(GA_SUB(1.0, d1), e1)));
    dq = GA_SUB(pow(v, mp), 1.0);
  } else {
    const double v = GA_ADD(GA_MUL(2.0, GA_SUB(1.0, r_u)),
                            GA_MUL(GA_MUL(2.0, GA_SUB(r_u, 0.5)), pow(GA_SUB(1.0, d2), e1)));
    dq = GA_SUB(1.0, pow(v, mp));
  }
  return clipd(GA_ADD(x, GA_MUL(dq, span)), xl, xu);
}

// Layout of the uniform draws of one call of the offspring operator (M matings of n_var variables), in the order
// the host operators consume them: SBX do / u / swap [M][V] each, SBX keep [M], PM do / u [2M][V] each.
GA_HD size_t rand_count(int M, int V) { return (size_t)7 * M * V + M; }

// bounds: [4][V] = operator lower / upper bound (for the integer variants already widened by 0.5 - 1e-16 as
// ga._IntegerFromFloat._Shift does), then the variable's own lower / upper bound used by the integer rounding clip.
GA_HD void offspring_element(const OpParams& p, const double* X, const int32_t* parents, const double* bounds,
                             const double* rnd, int M, int m, int v, double* out) {
  const int V = p.n_var;
  const size_t MV = (size_t)M * V, e = (size_t)m * V + v;
  const double x0 = X[(size_t)parents[2 * m] * V + v], x1 = X[(size_t)parents[2 * m + 1] * V + v];
  const double xl = bounds[v], xu = bounds[V + v];
  double c0, c1;
  sbx_element(p, x0, x1, xl, xu, rnd[e], rnd[MV + e], rnd[2 * MV + e], rnd[3 * MV + m], c0, c1);
  if (p.integer) {
    c0 = clipd(rint(c0), bounds[2 * V + v], bounds[3 * V + v]);
    c1 = clipd(rint(c1), bounds[2 * V + v], bounds[3 * V + v]);
  }
  const double* pm_do = rnd + 3 * MV + M;
  const double* pm_u = pm_do + 2 * MV;
  c0 = pm_element(p, c0, xl, xu, pm_do[e], pm_u[e]);
  c1 = pm_element(p, c1, xl, xu, pm_do[MV + e], pm_u[MV + e]);
  if (p.integer) {
    c0 = clipd(rint(c0), bounds[2 * V + v], bounds[3 * V + v]);
    c1 = clipd(rint(c1), bounds[2 * V + v], bounds[3 * V + v]);
  }
  out[e] = c0;                 // C.reshape(-1, V): first children of all matings, then the second children
  out[MV + e] = c1;
}

// Binary tournament (ga.Algorithm._tournament): pair t = (perm[2t], perm[2t+1]); the lower rank wins, then the
// larger crowding distance, then the first of the pair.
GA_HD int32_t tournament_element(const int32_t* pairs, const int32_t* rank, const double* crowd, int t) {
  const int32_t a = pairs[2 * t], b = pairs[2 * t + 1];
  const bool better_a = rank[a] < rank[b] || (rank[a] == rank[b] && crowd[a] >= crowd[b]);
  return better_a ? a : b;
}

// Position of element j in the stable ascending sort of key[0..n): the number of elements that sort before it.
GA_HD int stable_rank(const double* key, int n, int j) {
  const double kj = key[j];
  int r = 0;
  for (int i = 0; i < n; ++i) r += (key[i] < kj || (key[i] == kj && i < j)) ? 1 : 0;
  return r;
}

// Duplicate test (ga.Algorithm._mate): a candidate is dropped if it equals (max |difference| <= eps) a member of
// the population, an offspring accepted earlier, or an earlier candidate of this batch.
GA_HD bool rows_equal(const double* a, const double* b, int V, double eps) {
  for (int v = 0; v < V; ++v)
    if (!(fabs(GA_SUB(a[v], b[v])) <= eps)) return false;
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// Survival (ga.rank_and_crowding_survival / Algorithm._survive).  One cooperating group of threads (a CUDA block; on
// the host a plain loop: every GA_FOR runs to completion before the next one starts, which is what the barriers
// guarantee on the device).  F is column-major: objective k of candidate j at F[k * ld + j] (fp32, widened exactly).
// ---------------------------------------------------------------------------------------------------------------
struct SurviveState {
  const float* F;
  int ld, n, n_obj, n_survive;
  int nsga2;            // 1: fast non-dominated sort + crowding; 0: single-objective GA (sorted by F[0])
  // workspace, n entries each
  int* ndom;            // number of unassigned candidates that dominate j
  int* front;           // rank of j, -1 while unassigned
  int* cur;             // 1 if j is in the front being processed
  int* pos;             // position of j in the current per-objective sort
  double* cd;           // crowding distance of j
  double* sorted;       // objective values of the current front in ascending order
  int* fsize;           // one counter (shared by the group)
  // outputs, n_survive entries each
  int32_t* out_idx;
  int32_t* out_rank;
  double* out_crowd;
};

GA_HD bool dominates(const SurviveState& s, int i, int j) {
  bool le = true, lt = false;
  for (int k = 0; k < s.n_obj; ++k) {
    const float a = s.F[(size_t)k * s.ld + i], b = s.F[(size_t)k * s.ld + j];
    le = le && (a <= b);
    lt = lt || (a < b);
  }
  return le && lt;
}

GA_HD void survive_body(SurviveState& s) {
  const int n = s.n;
  if (!s.nsga2) {       // Algorithm._survive, "ga": idx = argsort(F[:, 0], stable)[:S]; rank 0; crowd = -F
    GA_FOR(j, n) {
      const float fj = s.F[j];
      int r = 0;
      for (int i = 0; i < n; ++i) {
        const float fi = s.F[i];
        r += (fi < fj || (fi == fj && i < j)) ? 1 : 0;
      }
      if (r < s.n_survive) {
        s.out_idx[r] = j;
        s.out_rank[r] = 0;
        s.out_crowd[r] = -(double)fj;
      }
    }
    GA_SYNC();
    return;
  }
  GA_FOR(j, n) {
    int c = 0;
    for (int i = 0; i < n; ++i) c += dominates(s, i, j) ? 1 : 0;
    s.ndom[j] = c;
    s.front[j] = -1;
  }
  if (GA_TID == 0) *s.fsize = 0;
  GA_SYNC();
  int cum = 0;
  for (int r = 0; r < n && cum < s.n_survive; ++r) {
    GA_FOR(j, n) {
      const int in = (s.front[j] < 0 && s.ndom[j] == 0) ? 1 : 0;
      s.cur[j] = in;
      if (in) GA_ATOMIC_ADD(s.fsize, 1);
    }
    GA_SYNC();
    const int fs = *s.fsize;
    if (fs == 0) break;                                   // every candidate is ranked
    // crowding distance inside the front (ga.crowding_distance on F[front], front in ascending index order)
    GA_FOR(j, n) if (s.cur[j]) s.cd[j] = 0.0;
    for (int k = 0; k < s.n_obj; ++k) {
      const float* Fk = s.F + (size_t)k * s.ld;
      GA_FOR(j, n) if (s.cur[j]) {
        const float fj = Fk[j];
        int p = 0;
        for (int i = 0; i < n; ++i)
          if (s.cur[i]) p += (Fk[i] < fj || (Fk[i] == fj && i < j)) ? 1 : 0;
        s.pos[j] = p;
        s.sorted[p] = (double)fj;
      }
      GA_SYNC();
      GA_FOR(j, n) if (s.cur[j]) {
        const int p = s.pos[j];
        if (fs <= 2 || p == 0 || p == fs - 1) {
          s.cd[j] = INFINITY;
        } else {
          const double span = GA_SUB(s.sorted[fs - 1], s.sorted[0]);
          if (span > 0) s.cd[j] = GA_ADD(s.cd[j], GA_DIV(GA_SUB(s.sorted[p + 1], s.sorted[p - 1]), span));
        }
      }
      GA_SYNC();
    }
    // survivors of this front: all of it in index order, or the n_survive - cum most isolated (stable sort by -cd)
    const int room = s.n_survive - cum;
    GA_FOR(j, n) if (s.cur[j]) {
      int p = 0;
      if (fs <= room) {
        for (int i = 0; i < j; ++i) p += s.cur[i];
      } else {
        const double cj = s.cd[j];
        for (int i = 0; i < n; ++i)
          if (s.cur[i]) p += (s.cd[i] > cj || (s.cd[i] == cj && i < j)) ? 1 : 0;
      }
      if (p < room) {
        s.out_idx[cum + p] = j;
        s.out_rank[cum + p] = r;
        s.out_crowd[cum + p] = s.cd[j];
      }
    }
    // peel the front
    GA_FOR(j, n) {
      if (s.cur[j]) {
        s.front[j] = r;
      } else if (s.front[j] < 0) {
        int c = 0;
        for (int i = 0; i < n; ++i)
          if (s.cur[i]) c += dominates(s, i, j) ? 1 : 0;
        s.ndom[j] -= c;
      }
    }
    GA_SYNC();
    if (GA_TID == 0) *s.fsize = 0;
    GA_SYNC();
    cum += fs;
  }
}

}  // namespace glass_ga
