// Fused per-structure kernel for padded batches (BASELINE configs 1-3, 5b).
//
// One CTA owns one structure of the batch and runs the whole hot path of the
// reference's `dftd3()` (disp.py:150-165) for it without touching HBM between
// stages and without any N x N (or N x N x 7 x 7) temporary:
//
//   phase 0  load Z / x / per-atom data, counting-sort the atoms by element
//   phase 1  CN sweep                      (tad-mctc cn_d3 + exp_count, r <= 25)
//   phase 2  Gaussian reference weights w, dw/dCN   (model/weights.py:104-176)
//   phase 3  pair sweep: C6_ij, dC6_ij/dCN_i, BJ energy, direct gradient,
//            dL/dCN_i, parameter gradients  (model/c6.py, disp.py:276-302,
//            damping/rational.py:68)
//   phase 3b ATM triple sweep (optional; damping/atm.py:86-155)
//   phase 4  CN-backward sweep (chain rule through the counting function)
//
// Thread t owns the atom at sorted position t and walks all partners j with a
// warp-uniform j, so every shared-memory read in the inner loops is a
// broadcast and nothing is scattered: no atomics, bitwise reproducible.
// Atoms are sorted by element so that the 7x7 reference block K[Zi,Zj] is
// contracted with the thread's own weights once per (atom, partner element):
//   U[b] = sum_a w_ia K[Zi,Zj,a,b],  Ud[b] = sum_a dw_ia K[Zi,Zj,a,b]
// after which C6_ij = U . w_j and dC6_ij/dCN_i = Ud . w_j cost 7 FMAs each
// (the reference gathers 49 values per pair, model/c6.py:216).
#pragma once
#include "d3_math.cuh"

namespace d3 {

constexpr int kMaxZ = 104;   // reference defaults.MAX_ELEMENT
constexpr int kNRef = 7;     // reference systems per element

enum : int { kWantE = 1, kWantG = 2, kWantP = 4, kAtm = 8 };

template <typename S>
struct MolArgs {
  using B = typename Traits<S>::base;
  int nbatch, natoms;
  const int64_t* numbers;   // [nbatch, natoms]
  const B* positions;       // [nbatch, natoms, 3]
  const B* dpositions;      // tangent of positions (Dual only) or null
  const B* rcov;            // per atom [nbatch, natoms] or per element [>=104]
  const B* r4r2;            // same convention
  int rcov_per_element, r4r2_per_element;
  const B* refcn;           // [104, 7]
  const B* refc6;           // [104, 104, 7, 7]
  const B* rvdw;            // mode 1: [104,104] table, mode 2: dense [nbatch,N,N]
  int rvdw_mode;
  const B* g;               // [nbatch, natoms] cotangent of the per-atom energy
  S s6, s8, a1, a2, s9, rs9, alp;
  B cutoff2, cn_cutoff2, kcn;
  B* energy;                // [nbatch, natoms]
  B* grad;                  // [nbatch, natoms, 3]
  B* pgrad;                 // [nbatch, 5]  (s6, s8, a1, a2, s9)
  B* denergy;               // tangents (Dual only)
  B* dgrad;
  B* dpgrad;
  S* work;                  // ATM scratch: sqrt|C6| per structure, [nbatch][N][N]
  const int* order;         // optional: CTA k works on structure order[k] (scheduling hint)
  int counting, damping;    // model functions (d3b200_params.counting / .damping); molecule_kernel only
  const int* ready;         // molecule_kernel_v2, streamed host form: structure b may be read once ready[b / ready_per] != 0
  int ready_per;            // (set by a stream-ordered memset behind the copy of each chunk of inputs)
  int host_io;              // molecule_kernel_v2: numbers / positions / energy / grad may live in mapped HOST
                            // memory -> every global access of these arrays fully coalesced
};

template <typename S> __device__ __forceinline__ S warp_sum(S v);
template <> __device__ __forceinline__ float warp_sum<float>(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <> __device__ __forceinline__ double warp_sum<double>(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <> __device__ __forceinline__ Dual<float> warp_sum<Dual<float>>(Dual<float> v) {
  return Dual<float>(warp_sum(v.v), warp_sum(v.d));
}
template <> __device__ __forceinline__ Dual<double> warp_sum<Dual<double>>(Dual<double> v) {
  return Dual<double>(warp_sum(v.v), warp_sum(v.d));
}

// shared memory bytes needed by the fused kernel for `natoms` atoms
template <typename S>
__host__ __device__ inline size_t molecule_smem_bytes(int natoms) {
  // S arrays: x y z rcov q sqrtq cn decn gx gy gz (11) + w[8] + dw[8] = 27 per atom
  // base array: g (1)   int arrays: origZ, sortedZ, perm (3)
  size_t n = (size_t)natoms;
  size_t bytes = n * 27 * sizeof(S) + n * sizeof(typename Traits<S>::base) + n * 3 * sizeof(int);
  bytes += (4 * kMaxZ + 8) * sizeof(int);          // cnt, zstart, gZ, gstart, counters
  bytes += 32 * 5 * sizeof(S);                      // block reduction scratch
  return bytes + 64;
}

// Gaussian weights of one atom and their CN derivative.
// Reference: model/weights.py:104-176 -- the CN difference is formed in the
// working precision, the exponentials and the normalisation in float64.
template <typename S>
__device__ __forceinline__ void reference_weights(
    const typename Traits<S>::base* __restrict__ refcn_row, S cn, S* w, S* dw) {
  using B = typename Traits<S>::base;
  using W = typename Traits<S>::wide;
  W gw[kNRef], dc[kNRef];
  W norm = W(0.0);
  B maxcn = B(-1e30);
  bool valid[kNRef];
#pragma unroll
  for (int a = 0; a < kNRef; ++a) {
    B rc = refcn_row[a];
    valid[a] = rc >= B(0);
    maxcn = rc > maxcn ? rc : maxcn;
    S diff = rc - cn;                 // working precision, as the reference
    dc[a] = widen(diff);
    gw[a] = valid[a] ? ex(-(4.0 * (dc[a] * dc[a]))) : W(0.0);
    norm += gw[a];
  }
  if (val(norm) == 0.0) {
    // every Gaussian underflowed: weight 1 on the reference(s) with the largest
    // reference CN, no CN dependence (weights.py:162-174)
#pragma unroll
    for (int a = 0; a < kNRef; ++a) {
      w[a] = (valid[a] && refcn_row[a] == maxcn) ? S(B(1)) : S(B(0));
      dw[a] = S(B(0));
    }
    return;
  }
  // true division: when CN is far from every reference CN the Gaussians and
  // their sum are denormal (1e-310); 1/norm would overflow where g/norm is exact
  W wa[kNRef];
  W mean = W(0.0);
#pragma unroll
  for (int a = 0; a < kNRef; ++a) {
    wa[a] = dv(gw[a], norm);
    mean += wa[a] * dc[a];
  }
#pragma unroll
  for (int a = 0; a < kNRef; ++a) {
    // d w_a / d CN = 8 w_a (dc_a - sum_b w_b dc_b),  dc = refcn - CN
    W d = 8.0 * (wa[a] * (dc[a] - mean));
    w[a] = narrow<S>(wa[a]);
    dw[a] = narrow<S>(d);
  }
}

// ---- model functions behind the reference's callable hooks (disp.py:89-91) ---------------------------
enum : int { kCountExp = 0, kCountErf = 1 };
enum : int { kDampRational = 0, kDampZero = 1 };

// Counting function f(r) and df = f'(r) / r of one pair (r2 = r^2 > 0, rc = rcov_i + rcov_j).
//   exp (D3, tad-mctc exp_count):  f = 1 / (1 + exp(-kcn (rc / r - 1)))
//   erf (tad-mctc erf_count):      f = (1 + erf(-kcn (r / rc - 1))) / 2
template <typename S, typename B>
__device__ __forceinline__ void counting_pair(int kind, S r2, S rc, B kcn, S& f, S& df) {
  S rinv = frsq(r2);
  if (kind == kCountErf) {
    S x = r2 * rinv * rcp(rc);                    // r / rc
    S y = kcn * (B(1) - x);
    f = B(0.5) * (B(1) + er(y));
    // f' = -kcn / (rc sqrt(pi)) exp(-y^2)
    df = (B(-0.5641895835477563) * kcn) * (rcp(rc) * rinv) * ex(-(y * y));
    return;
  }
  S arg = kcn * (rc * rinv - B(1));
  S e = fex(fmx(-arg, B(-700)));
  f = frcp(B(1) + e);
  df = (-kcn) * (rc * (rinv * rinv * rinv)) * (f * (e * f));   // f (1 - f) = e f^2
}

// Two-body radial factors t6, t8 (E_ij = -C6 (s6 t6 + s8 qq t8)), their derivatives with respect to r^2
// and with respect to the two damping parameters (p1, p2) = (a1, a2) or (rs6, rs8).
//   rational (BJ, damping/rational.py:68):  t_n = 1 / (r^n + R0^n),  R0 = a1 sqrt(qq) + a2
//   zero (D3(0)-style on R0 = sqrt(qq)):     t_n = r^-n / (1 + 6 (rs_n R0 / r)^alp_n)
template <typename S, typename B>
__device__ __forceinline__ void damping_pair(int kind, S r2, S sqq, S p1, S p2, B alp6,
                                             S& t6, S& t8, S& dt6, S& dt8,     // dt_n = -d t_n / d(r^2)
                                             S& t6p1, S& t6p2, S& t8p1, S& t8p2) {
  if (kind == kDampZero) {
    const B alp8 = alp6 + B(2);
    S rinv = frsq(r2);
    S ir2 = rinv * rinv;
    S ir6 = ir2 * ir2 * ir2;
    S u6 = pw(p1 * (sqq * rinv), alp6), u8 = pw(p2 * (sqq * rinv), alp8);
    S f6 = frcp(B(1) + B(6) * u6), f8 = frcp(B(1) + B(6) * u8);
    t6 = ir6 * f6;
    t8 = (ir6 * ir2) * f8;
    // d t_n / d r^2 = t_n / r^2 (-n/2 + 3 alp_n u_n f_n)
    dt6 = (t6 * ir2) * (B(3) - (B(3) * alp6) * (u6 * f6));
    dt8 = (t8 * ir2) * (B(4) - (B(3) * alp8) * (u8 * f8));
    // d t_n / d rs_n = -6 alp_n t_n f_n u_n / rs_n
    t6p1 = (B(-6) * alp6) * ((t6 * f6) * (u6 * rcp(p1)));
    t8p2 = (B(-6) * alp8) * ((t8 * f8) * (u8 * rcp(p2)));
    t6p2 = S(B(0));
    t8p1 = S(B(0));
    return;
  }
  S r0 = p1 * sqq + p2;
  S r02 = r0 * r0;
  S r06 = r02 * r02 * r02;
  S r08 = r06 * r02;
  S r4 = r2 * r2;
  S r6 = r4 * r2;
  S r8 = r4 * r4;
  t6 = frcp(r6 + r06);
  t8 = frcp(r8 + r08);
  S t62 = t6 * t6, t82 = t8 * t8;
  dt6 = B(3) * (r4 * t62);
  dt8 = B(4) * (r6 * t82);
  // d t_n / d R0 = -n R0^(n-1) t_n^2;  dR0/da1 = sqrt(qq), dR0/da2 = 1
  S r05 = r02 * r02 * r0;
  t6p2 = B(-6) * (r05 * t62);
  t8p2 = B(-8) * ((r05 * r02) * t82);
  t6p1 = t6p2 * sqq;
  t8p1 = t8p2 * sqq;
}

// One ATM term and its derivative with respect to a = r_ij^2
// (reference damping/atm.py:93-153).
//   a, b, c : r_ij^2, r_ik^2, r_jk^2      inva = 1/a
//   c9      : s9 sqrt|C6_ij C6_ik C6_jk|  r0 = rs9^3 R_ij R_ik R_jk
//   pexp    : (alp + 2) / 3
template <typename S, typename B>
__device__ __forceinline__ void atm_term(S a, S b, S c, S inva, S c9, S r0, B pexp, S& e, S& de_da) {
  S r2 = a * b * c;
  S r1inv = frsq(r2);
  S r3inv = r1inv * r1inv * r1inv;
  S r5inv = r3inv * (r1inv * r1inv);
  S p = a + c - b, q = a - c + b, t = (c + b) - a;
  S s = p * q * t;
  S ang = B(0.375) * (s * r5inv) + r3inv;
  S u = pw(r0 * r1inv, pexp);
  S fd = frcp(B(1) + B(6) * u);
  e = c9 * (ang * fd);
  S ds = q * t + p * t - p * q;
  S dang = B(0.375) * (ds * r5inv - B(2.5) * ((s * r5inv) * inva)) - B(1.5) * (r3inv * inva);
  S dfd = (B(3) * pexp) * (u * (fd * fd) * inva);
  de_da = c9 * (ang * dfd + fd * dang);
}

constexpr int kMolThreads = 256;  // CTA size cap; larger structures loop over i

template <typename S, int FLAGS>
__global__ void __launch_bounds__(kMolThreads) molecule_kernel(MolArgs<S> A) {
  using B = typename Traits<S>::base;
  constexpr bool DUAL = Traits<S>::dual;
  constexpr bool WANT_E = FLAGS & kWantE;
  constexpr bool WANT_G = FLAGS & kWantG;
  constexpr bool WANT_P = FLAGS & kWantP;
  constexpr bool ATM = FLAGS & kAtm;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = A.natoms;
  const int b = A.order ? A.order[blockIdx.x] : blockIdx.x;
  const int t = threadIdx.x;
  const int nt = blockDim.x;

  S* sx = reinterpret_cast<S*>(smem_raw);
  S* sy = sx + N;
  S* sz = sy + N;
  S* srcov = sz + N;
  S* sq = srcov + N;
  S* ssq = sq + N;
  S* scn = ssq + N;
  S* sdecn = scn + N;
  S* sgx = sdecn + N;
  S* sgy = sgx + N;
  S* sgz = sgy + N;
  S* sw = sgz + N;              // [N][8]
  S* sdw = sw + 8 * (size_t)N;  // [N][8]
  S* sred = sdw + 8 * (size_t)N;  // [32][5]
  B* sg = reinterpret_cast<B*>(sred + 32 * 5);
  int* sorigZ = reinterpret_cast<int*>(sg + N);
  int* sZ = sorigZ + N;
  int* sperm = sZ + N;
  int* cnt = sperm + N;
  int* zstart = cnt + kMaxZ;
  int* gZ = zstart + kMaxZ;
  int* gstart = gZ + kMaxZ;     // [kMaxZ + 1]
  int* misc = gstart + kMaxZ + 1;  // [0] = ngroups, [1] = nreal

  const int64_t* numbers = A.numbers + (size_t)b * N;
  const B* pos = A.positions + (size_t)b * N * 3;

  // ---------------------------------------------------------------- phase 0
  for (int k = t; k < kMaxZ; k += nt) cnt[k] = 0;
  __syncthreads();
  for (int k = t; k < N; k += nt) {
    int Z = (int)numbers[k];
    if (Z < 0 || Z >= kMaxZ) Z = 0;  // host validates; never index out of range
    sorigZ[k] = Z;
    if (Z > 0) {
      atomicAdd(&cnt[Z], 1);
    } else {
      // padding / ghost atom: exact zeros in every output (reference: masks)
      size_t o = (size_t)b * N + k;
      if (WANT_E && A.energy) A.energy[o] = B(0);
      if (WANT_E && DUAL && A.denergy) A.denergy[o] = B(0);
      if (WANT_G) {
        if (A.grad) { A.grad[3 * o] = B(0); A.grad[3 * o + 1] = B(0); A.grad[3 * o + 2] = B(0); }
        if (DUAL && A.dgrad) { A.dgrad[3 * o] = B(0); A.dgrad[3 * o + 1] = B(0); A.dgrad[3 * o + 2] = B(0); }
      }
    }
  }
  __syncthreads();
  if (t == 0) {
    int ng = 0, run = 0;
    for (int z = 1; z < kMaxZ; ++z) {
      zstart[z] = run;
      if (cnt[z] > 0) { gZ[ng] = z; gstart[ng] = run; ++ng; run += cnt[z]; }
    }
    gstart[ng] = run;
    misc[0] = ng;
    misc[1] = run;
  }
  __syncthreads();
  const int ng = misc[0];
  const int n = misc[1];
  for (int k = t; k < N; k += nt) {
    const int Z = sorigZ[k];
    if (Z == 0) continue;
    int p = zstart[Z];
    for (int m = 0; m < k; ++m) p += (sorigZ[m] == Z);
    sperm[p] = k;
    sZ[p] = Z;
    B px = pos[3 * k], py = pos[3 * k + 1], pz = pos[3 * k + 2];
    B dx = B(0), dy = B(0), dz = B(0);
    if (DUAL && A.dpositions) {
      const B* dp = A.dpositions + ((size_t)b * N + k) * 3;
      dx = dp[0]; dy = dp[1]; dz = dp[2];
    }
    sx[p] = make<S, B>(px, dx);
    sy[p] = make<S, B>(py, dy);
    sz[p] = make<S, B>(pz, dz);
    B rc = A.rcov_per_element ? A.rcov[Z] : A.rcov[(size_t)b * N + k];
    B q = A.r4r2_per_element ? A.r4r2[Z] : A.r4r2[(size_t)b * N + k];
    srcov[p] = S(rc);
    sq[p] = S(q);
    ssq[p] = S(sqr(q));
    sg[p] = (WANT_G && A.g) ? A.g[(size_t)b * N + k] : B(1);
  }
  __syncthreads();

  const B eps = eps_of<B>();
  S* hrow = ATM ? A.work + (size_t)b * N * N : nullptr;   // h[j*N + i] = sqrt|C6_ij|

  // ---------------------------------------------------------------- phase 1+2
  // CN_i = sum_j [r_ij <= 25] / (1 + exp(-kcn (rc_ij / r_ij - 1))), then weights
  for (int i = t; i < n; i += nt) {
    const S xi = sx[i], yi = sy[i], zi = sz[i], rci = srcov[i];
    S cn = S(B(0));
    for (int j = 0; j < n; ++j) {
      if (j == i) continue;
      S dx = xi - sx[j], dy = yi - sy[j], dz = zi - sz[j];
      S r2 = dx * dx + dy * dy + dz * dz;
      if (val(r2) > A.cn_cutoff2) continue;
      if (val(r2) < eps) r2 = S(eps);
      if (A.counting == kCountExp) {
        S rinv = frsq(r2);
        S arg = A.kcn * ((rci + srcov[j]) * rinv - B(1));
        cn += frcp(B(1) + fex(fmx(-arg, B(-700))));
      } else {
        S f, df;
        counting_pair<S, B>(A.counting, r2, rci + srcov[j], A.kcn, f, df);
        cn += f;
      }
    }
    scn[i] = cn;
    S wi[kNRef], dwi[kNRef];
    reference_weights<S>(A.refcn + sZ[i] * kNRef, cn, wi, dwi);
#pragma unroll
    for (int a = 0; a < kNRef; ++a) { sw[8 * i + a] = wi[a]; sdw[8 * i + a] = dwi[a]; }
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase 3
  S p_s6 = S(B(0)), p_s8 = S(B(0)), p_a1 = S(B(0)), p_a2 = S(B(0));
  for (int i = t; i < n; i += nt) {
    const S xi = sx[i], yi = sy[i], zi = sz[i], qi = sq[i];
    const int Zi = sZ[i];
    S wi[kNRef], dwi[kNRef];
#pragma unroll
    for (int a = 0; a < kNRef; ++a) { wi[a] = sw[8 * i + a]; dwi[a] = sdw[8 * i + a]; }
    S e_i = S(B(0)), gx = S(B(0)), gy = S(B(0)), gz = S(B(0)), decn = S(B(0));
    const B gi = sg[i];
    const S s3q_i = sqr(B(3) * qi);         // sqrt(3 q_i);  sqrt(qq) = s3q_i sqrt(q_j)
    const S q3_i = B(3) * qi;
    for (int grp = 0; grp < ng; ++grp) {
      const int Zj = gZ[grp];
      const B* __restrict__ K = A.refc6 + ((size_t)Zi * kMaxZ + Zj) * (kNRef * kNRef);
      S U[kNRef], Ud[kNRef];
#pragma unroll
      for (int bb = 0; bb < kNRef; ++bb) { U[bb] = S(B(0)); Ud[bb] = S(B(0)); }
#pragma unroll
      for (int a = 0; a < kNRef; ++a) {
#pragma unroll
        for (int bb = 0; bb < kNRef; ++bb) {
          B k = __ldg(K + a * kNRef + bb);
          U[bb] += wi[a] * k;
          if (WANT_G) Ud[bb] += dwi[a] * k;
        }
      }
      const int j1 = gstart[grp + 1];
      for (int j = gstart[grp]; j < j1; ++j) {
        if (j == i) continue;
        S dx = xi - sx[j], dy = yi - sy[j], dz = zi - sz[j];
        S r2 = dx * dx + dy * dy + dz * dz;
        if (!ATM && val(r2) > A.cutoff2) continue;
        const S* wj = sw + 8 * j;
        S c6 = U[0] * wj[0];
#pragma unroll
        for (int a = 1; a < kNRef; ++a) c6 += U[a] * wj[a];
        if (ATM) {
          // the three-body term needs C6 of every pair (r_ik is not cut off)
          hrow[(size_t)j * N + i] = sqr(ab(c6));
          if (val(r2) > A.cutoff2) continue;
        }
        if (val(r2) < eps) r2 = S(eps);
        S qq = q3_i * sq[j];
        S sqq = s3q_i * ssq[j];                    // sqrt(qq) = sqrt(3 q_i q_j)
        S t6, t8, dt6, dt8, t6p1, t6p2, t8p1, t8p2;
        damping_pair<S, B>(A.damping, r2, sqq, A.a1, A.a2, val(A.alp), t6, t8, dt6, dt8, t6p1, t6p2, t8p1, t8p2);
        S s8q = A.s8 * qq;
        S P = A.s6 * t6 + s8q * t8;
        if (WANT_E) e_i += c6 * P;   // scaled by -1/2 at the end
        if (WANT_G) {
          S dc6 = Ud[0] * wj[0];
#pragma unroll
          for (int a = 1; a < kNRef; ++a) dc6 += Ud[a] * wj[a];
          B gam = B(0.5) * (gi + sg[j]);
          // dL/dC6_ij = -gam P ;  dL/dCN_i += that * dC6_ij/dCN_i
          decn -= (gam * P) * dc6;
          // -dP/d(r^2) = s6 dt6 + s8 qq dt8
          S dP = A.s6 * dt6 + s8q * dt8;
          // dL/dx_i = -gam C6 dP/dr2 * 2 d
          S f = (B(2) * gam) * (c6 * dP);
          gx += f * dx; gy += f * dy; gz += f * dz;
          if (WANT_P) {
            // each unordered pair is visited from both ends: weight 1/2
            S hc = (B(-0.5) * gam) * c6;
            p_s6 += hc * t6;
            p_s8 += hc * (qq * t8);
            // derivatives with respect to the two damping parameters (a1, a2) / (rs6, rs8)
            p_a1 += hc * (A.s6 * t6p1 + s8q * t8p1);
            p_a2 += hc * (A.s6 * t6p2 + s8q * t8p2);
          }
        }
      }
    }
    if (WANT_E) {
      size_t o = (size_t)b * N + sperm[i];
      S e = B(-0.5) * e_i;
      if (A.energy) A.energy[o] = val(e);
      if (DUAL && A.denergy) A.denergy[o] = tan_of(e);
    }
    if (WANT_G) { sdecn[i] = decn; sgx[i] = gx; sgy[i] = gy; sgz[i] = gz; }
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase 3b
  // Axilrod-Teller-Muto term (damping/atm.py:86-155) in ordered form:
  //   E_i = 1/6 sum_j sum_k [r_ij<=c][r_jk<=c] e_ijk        (no test on r_ik)
  // Thread i walks ordered (j, k); e_ijk is symmetric, so with
  //   w_ijk = g_i m_jk(m_ij+m_ik) + g_j m_ik(m_ij+m_jk) + g_k m_ij(m_ik+m_jk)
  //   dL/dx_i   = 1/6 sum_j [ sum_k w de/d(r_ij^2) ] 2 (x_i - x_j)
  //   dL/dCN_i += 1/6 sum_j [ sum_k w e ] / (2 C6_ij) dC6_ij/dCN_i
  // everything accumulates in registers of the owning thread.
  S p_s9 = S(B(0));
  if (ATM) {
    const B pexp = (val(A.alp) + B(2)) / B(3);
    const S rs93 = A.rs9 * A.rs9 * A.rs9;
    const B sixth = B(1) / B(6);
    for (int i = t; i < n; i += nt) {
      const S xi = sx[i], yi = sy[i], zi = sz[i];
      const int Zi = sZ[i];
      const int oi = sperm[i];
      const B gi = sg[i];
      S wi[kNRef], dwi[kNRef];
#pragma unroll
      for (int a = 0; a < kNRef; ++a) { wi[a] = sw[8 * i + a]; dwi[a] = sdw[8 * i + a]; }
      S e_atm = S(B(0)), gx = S(B(0)), gy = S(B(0)), gz = S(B(0)), decn = S(B(0));
      for (int gj = 0; gj < ng; ++gj) {
        const int Zj = gZ[gj];
        S U[kNRef], Ud[kNRef];
        if (WANT_G) {
          const B* __restrict__ K = A.refc6 + ((size_t)Zi * kMaxZ + Zj) * (kNRef * kNRef);
#pragma unroll
          for (int bb = 0; bb < kNRef; ++bb) { U[bb] = S(B(0)); Ud[bb] = S(B(0)); }
#pragma unroll
          for (int a = 0; a < kNRef; ++a) {
#pragma unroll
            for (int bb = 0; bb < kNRef; ++bb) {
              B k = __ldg(K + a * kNRef + bb);
              U[bb] += wi[a] * k;
              Ud[bb] += dwi[a] * k;
            }
          }
        }
        const B rij_tab = A.rvdw_mode == 1 ? A.rvdw[Zi * kMaxZ + Zj] : B(0);
        for (int j = gstart[gj]; j < gstart[gj + 1]; ++j) {
          if (j == i) continue;
          const S xj = sx[j], yj = sy[j], zj = sz[j];
          S dx = xi - xj, dy = yi - yj, dz = zi - zj;
          S a = dx * dx + dy * dy + dz * dz;
          if (val(a) < eps) a = S(eps);
          const bool mij = val(a) <= A.cutoff2;
          const S inva = frcp(a);
          const S hij = hrow[(size_t)j * N + i];
          const int oj = sperm[j];
          const B gj_ = sg[j];
          const B Rij = A.rvdw_mode == 1 ? rij_tab : A.rvdw[((size_t)b * N + oi) * N + oj];
          S Fj = S(B(0)), Gj = S(B(0)), Ej = S(B(0));
          for (int gk = 0; gk < ng; ++gk) {
            const int Zk = gZ[gk];
            const B rik_tab = A.rvdw_mode == 1 ? A.rvdw[Zi * kMaxZ + Zk] : B(0);
            const B rjk_tab = A.rvdw_mode == 1 ? A.rvdw[Zj * kMaxZ + Zk] : B(0);
            for (int k = gstart[gk]; k < gstart[gk + 1]; ++k) {
              if (k == i || k == j) continue;
              const S xk = sx[k], yk = sy[k], zk = sz[k];
              S ex_ = xi - xk, ey = yi - yk, ez = zi - zk;
              S bq = ex_ * ex_ + ey * ey + ez * ez;
              S fx = xj - xk, fy = yj - yk, fz = zj - zk;
              S cq = fx * fx + fy * fy + fz * fz;
              if (val(bq) < eps) bq = S(eps);
              if (val(cq) < eps) cq = S(eps);
              const bool mik = val(bq) <= A.cutoff2, mjk = val(cq) <= A.cutoff2;
              if ((int)mij + (int)mik + (int)mjk < 2) continue;
              B Rik, Rjk;
              if (A.rvdw_mode == 1) { Rik = rik_tab; Rjk = rjk_tab; }
              else {
                const int ok = sperm[k];
                Rik = A.rvdw[((size_t)b * N + oi) * N + ok];
                Rjk = A.rvdw[((size_t)b * N + oj) * N + ok];
              }
              S c9 = A.s9 * (hij * (hrow[(size_t)k * N + i] * hrow[(size_t)k * N + j]));
              S r0 = rs93 * (Rij * Rik * Rjk);
              S e, de;
              atm_term<S, B>(a, bq, cq, inva, c9, r0, pexp, e, de);
              if (mij && mjk) Ej += e;
              if (WANT_G) {
                B w = gi * (mjk ? B((int)mij + (int)mik) : B(0))
                    + gj_ * (mik ? B((int)mij + (int)mjk) : B(0))
                    + sg[k] * (mij ? B((int)mik + (int)mjk) : B(0));
                Fj += w * de;
                Gj += w * e;
              }
            }
          }
          e_atm += Ej;
          if (WANT_G) {
            S f = (B(2) * sixth) * Fj;
            gx += f * dx; gy += f * dy; gz += f * dz;
            // dC6_ij/dCN_i / (2 C6_ij);  C6_ij = U . w_j
            const S* wj = sw + 8 * j;
            S c6 = U[0] * wj[0], dc6 = Ud[0] * wj[0];
#pragma unroll
            for (int aa = 1; aa < kNRef; ++aa) { c6 += U[aa] * wj[aa]; dc6 += Ud[aa] * wj[aa]; }
            decn += (B(0.5) * sixth) * (Gj * (dc6 * rcp0(c6)));
          }
        }
      }
      S e3 = sixth * e_atm;
      if (WANT_E) {
        size_t o = (size_t)b * N + oi;
        if (A.energy) A.energy[o] += val(e3);
        if (DUAL && A.denergy) A.denergy[o] += tan_of(e3);
      }
      if (WANT_G) {
        sdecn[i] += decn; sgx[i] += gx; sgy[i] += gy; sgz[i] += gz;
        if (WANT_P) p_s9 += gi * (e3 * rcp(A.s9));
      }
    }
    __syncthreads();
  }

  // ---------------------------------------------------------------- phase 4
  // dL/dx_i += sum_j (dL/dCN_i + dL/dCN_j) df_ij/dx_i,
  // df/dx_i = -kcn rc r^-3 f (1 - f) (x_i - x_j)
  if (WANT_G) {
    for (int i = t; i < n; i += nt) {
      const S xi = sx[i], yi = sy[i], zi = sz[i], rci = srcov[i], decn = sdecn[i];
      S gx = sgx[i], gy = sgy[i], gz = sgz[i];
      for (int j = 0; j < n; ++j) {
        if (j == i) continue;
        S dx = xi - sx[j], dy = yi - sy[j], dz = zi - sz[j];
        S r2 = dx * dx + dy * dy + dz * dz;
        if (val(r2) > A.cn_cutoff2) continue;
        if (val(r2) < eps) continue;  // clamped distance: no position dependence
        S f, df;
        counting_pair<S, B>(A.counting, r2, rci + srcov[j], A.kcn, f, df);
        S c = (decn + sdecn[j]) * df;
        gx += c * dx; gy += c * dy; gz += c * dz;
      }
      size_t o = (size_t)b * N + sperm[i];
      if (A.grad) { A.grad[3 * o] = val(gx); A.grad[3 * o + 1] = val(gy); A.grad[3 * o + 2] = val(gz); }
      if (DUAL && A.dgrad) { A.dgrad[3 * o] = tan_of(gx); A.dgrad[3 * o + 1] = tan_of(gy); A.dgrad[3 * o + 2] = tan_of(gz); }
    }
  }

  if (WANT_P) {
    S v[5] = {p_s6, p_s8, p_a1, p_a2, p_s9};
    const int lane = t & 31, wid = t >> 5, nw = (nt + 31) >> 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      S s = warp_sum<S>(v[k]);
      if (lane == 0) sred[wid * 5 + k] = s;
    }
    __syncthreads();
    if (t < 5) {
      S s = S(B(0));
      for (int w = 0; w < nw; ++w) s += sred[w * 5 + t];
      if (A.pgrad) A.pgrad[(size_t)b * 5 + t] = val(s);
      if (DUAL && A.dpgrad) A.dpgrad[(size_t)b * 5 + t] = tan_of(s);
    }
  }
}

}  // namespace d3
