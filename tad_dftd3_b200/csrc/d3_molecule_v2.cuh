// Fused per-structure kernel, second generation (the bench / production path
// for energies + first derivatives in float/double, two-body + CN).
//
// Same phases as `molecule_kernel` (d3_molecule.cuh) but every unordered pair
// is evaluated ONCE (Newton's third law) instead of from both ends:
//
//  * rows are processed in blocks of 16; a warp's two half-warps own the same
//    16 rows i and split the partner range j > i by parity, so j is uniform per
//    half-warp (shared-memory reads are broadcasts) and the triangular waste is
//    that of 16-row blocks (~85 % lane efficiency at N = 85);
//  * the strict upper triangle, linearised row-block by row-block, is cut into
//    four equal contiguous ranges, one per warp (fixed CTA of 4 warps): equal
//    work per warp, and a summation order that depends on the structure only;
//  * i-side contributions stay in registers; the j-side contributions of the 16
//    lanes are transposed through a small shared-memory tile, summed in a fixed
//    order by V lanes and added to per-warp private accumulators: no atomics,
//    bitwise reproducible, ~3x fewer instructions than a shuffle tree;
//  * per-atom data are AoS in shared memory ({x,y,z,kcn*rcov}, {sqrt q, g},
//    w[8], dw[8]) so the broadcast reads are 128-bit;
//  * rsqrt / reciprocal / exp are the branch-free MUFU-seeded versions of
//    d3_math.cuh, the two BJ reciprocals share one MUFU seed, and the sweeps
//    process several partners per iteration for instruction-level parallelism.
//
// Third generation (round 2), three changes that cut the instructions per pair-term by ~25 %:
//  * TMEM as a per-lane cache (d3_tmem.cuh): the CN sweep keeps kcn rc r^-3 f (1 - f) of every pair in the
//    SM's otherwise idle tensor memory and the CN-backward sweep reads it back instead of recomputing
//    rsqrt + exp + reciprocal (a lane re-reads what it wrote itself; both sweeps walk the pairs in the same order);
//  * reference-count trimming: an element with nref < 7 reference systems (H 2, C 5, N 4, O 3, ...) has exact
//    zeros in w[nref..6], so the three 7-wide dots per pair and the 7x7 contraction per (row block, partner
//    element) are instantiated for 2, 3, 5 and 7 columns and dispatched per partner element (warp-uniform);
//  * the validity of a partner (row exists, i < j < j1) is one unsigned compare feeding the cutoff compare.
#pragma once
#include <type_traits>

#include "d3_molecule.cuh"
#include "d3_tmem.cuh"

#ifndef D3_V2_TMEM
#define D3_V2_TMEM 1
#endif
#ifndef D3_V2_NREF
#define D3_V2_NREF 1
#endif

namespace d3 {

__device__ __forceinline__ float shx(float v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ double shx(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }

template <typename S> struct alignas(4 * sizeof(S)) Atom4 { S x, y, z, rk; };
template <typename S> struct alignas(2 * sizeof(S)) Pair2 { S a, b; };

constexpr int kV2Warps = 4;      // fixed CTA size (see above)
constexpr int kRedV = 10;        // values per transposed reduction and half-warp
constexpr int kRedStride = 18;   // 16 lanes + padding, keeps rows 16 B aligned
__host__ __device__ inline int v2_warps(int) { return kV2Warps; }

template <typename S>
__host__ __device__ inline size_t molecule_v2_smem_bytes(int natoms) {
  size_t n = (size_t)natoms;
  size_t bytes = n * (4 + 2 + 16) * sizeof(S);                // Atom4, Pair2 (later dL/dCN), w[8], dw[8]
  bytes += n * 5 * kV2Warps * sizeof(S);                      // per-warp accumulators [nw][5][N]
  bytes += (size_t)kV2Warps * 2 * kRedV * kRedStride * sizeof(S);  // reduction tiles
  bytes += n * 5 + 16;                                        // origZ, sortedZ, group (u8), perm (u16)
  bytes += (5 * kMaxZ + 8) * sizeof(int);
  return bytes + 128;
}

// lanes that share one row of the reduction tile (2V rows, 32 lanes)
__host__ __device__ constexpr int red_q(int V) { return 8 * V <= 32 ? 4 : (4 * V <= 32 ? 2 : 1); }

// Sum vals[v] over the 16 lanes of each half-warp (both halves at once).
// The values are transposed through this warp's tile [2][kRedV][kRedStride];
// Q = red_q(V) lanes then share each row r = half * V + v (a shuffle or two
// completes the row).  On return lanes with L % Q == 0 and L / Q < 2V hold the
// total of row L / Q.  The summation order is fixed.
template <int V, typename S>
__device__ __forceinline__ S rows_reduce(const S (&vals)[V], S* redw, int lane) {
  static_assert(V <= kRedV && 2 * V <= 32, "reduction tile too small");
  const int hl = lane & 15, ph = lane >> 4;
  S* mine = redw + ph * (kRedV * kRedStride);
#pragma unroll
  for (int v = 0; v < V; ++v) mine[v * kRedStride + hl] = vals[v];
  __syncwarp();
  S tot = S(0);
  constexpr int Q = red_q(V);   // lanes per row
  {
    const int r = lane / Q, part = lane % Q;
    if (r < 2 * V) {
      const S* row = redw + (r / V) * (kRedV * kRedStride) + (r % V) * kRedStride + part * (16 / Q);
      const Pair2<S>* rp = reinterpret_cast<const Pair2<S>*>(row);
      S s0 = S(0), s1 = S(0);
#pragma unroll
      for (int k = 0; k < 8 / Q; ++k) { Pair2<S> p = rp[k]; s0 += p.a; s1 += p.b; }
      tot = s0 + s1;
    }
  }
  if (Q >= 2) tot += shx(tot, 1);
  if (Q == 4) tot += shx(tot, 2);
  __syncwarp();
  return tot;
}
// decode helpers for the lanes that hold totals
template <int V> __device__ __forceinline__ bool red_owner(int lane) {
  constexpr int Q = red_q(V);
  return (lane % Q) == 0 && (lane / Q) < 2 * V;
}
template <int V> __device__ __forceinline__ int red_row(int lane) {
  constexpr int Q = red_q(V);
  return lane / Q;
}

// -(1/r) d f / d r = kcn rc r^-3 f (1 - f) of the counting function for one pair (0 when the pair does not
// count): what the CN sweep stores in TMEM, recomputed for the pairs that did not fit.
template <typename S>
__device__ __noinline__ S cn_df_recompute(S dx, S dy, S dz, S rk, bool pj, S cn_cutoff2, S kcn) {
  const S r2 = dx * dx + dy * dy + dz * dz;
  const bool ok = pj & (r2 <= cn_cutoff2) & (r2 >= eps_of<S>());
  const S rinv = frsq(r2 + tiny_of<S>());
  const S e = fex(fmx(kcn - rk * rinv, S(-700)));
  const S f = frcp(S(1) + e);
  const S dfv = (rk * (rinv * rinv * rinv)) * (f * (e * f));
  return ok ? dfv : S(0);
}

// QE: r4r2 comes from the per-element table (the default `r4r2=None` of dftd3()), so
// R0 = a1 sqrt(3 q_i q_j) + a2, R0^6, R0^8 and s8 * 3 q_i q_j depend on (row, partner
// ELEMENT) only and are hoisted out of the partner loop next to the U contraction.
template <typename S, bool WANT_G, bool QE>
__global__ void __launch_bounds__(32 * kV2Warps, 4) molecule_kernel_v2(MolArgs<S> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = A.natoms;
  const int b = A.order ? A.order[blockIdx.x] : blockIdx.x;
  const int t = threadIdx.x;
  constexpr int nt = 32 * kV2Warps;
  constexpr int nw = kV2Warps;
  const int lane = t & 31, wid = t >> 5;
  const int hl = lane & 15, ph = lane >> 4;

  Atom4<S>* sA = reinterpret_cast<Atom4<S>*>(smem_raw);
  Pair2<S>* sB = reinterpret_cast<Pair2<S>*>(sA + N);            // {sqrt(q), g}
  S* sw = reinterpret_cast<S*>(sB + N);                          // [N][8]
  S* sdw = sw + 8 * (size_t)N;                                   // [N][8]
  S* red = sdw + 8 * (size_t)N;                                  // [nw][2][kRedV][kRedStride]
  S* acc = red + (size_t)nw * 2 * kRedV * kRedStride;            // [nw][5][N]
  S* sdecn = reinterpret_cast<S*>(sB);                           // aliases sB after phase 3
  int* cnt = reinterpret_cast<int*>(acc + (size_t)nw * 5 * N);
  int* zstart = cnt + kMaxZ;
  int* gZ = zstart + kMaxZ;
  int* gstart = gZ + kMaxZ;
  int* zgrp = gstart + kMaxZ + 1;
  int* misc = zgrp + kMaxZ;
  unsigned short* sperm = reinterpret_cast<unsigned short*>(misc + 4);
  unsigned char* sorigZ = reinterpret_cast<unsigned char*>(sperm + N + (N & 1));
  unsigned char* sZ = sorigZ + N;
  unsigned char* sgrp = sZ + N;

  const int64_t* numbers = A.numbers + (size_t)b * N;
  const S* pos = A.positions + (size_t)b * N * 3;
  if (A.ready) {
    // streamed host form: the inputs of this structure are still on their way over PCIe (copy engine, chunk by
    // chunk, on another stream); wait for the flag that the copy stream sets behind the chunk
    if (t == 0) {
      const volatile int* flag = A.ready + b / A.ready_per;
      while (*flag == 0) __nanosleep(256);
      __threadfence();
    }
    __syncthreads();
  }

  // tensor memory: kCols 32-bit columns for this CTA, a quarter of the lanes per warp
  constexpr bool kTmem = WANT_G && (D3_V2_TMEM != 0);
  constexpr int kWordsPerVal = (int)(sizeof(S) / 4);
  constexpr int kTmemSlots = tmem::kCols / kWordsPerVal;   // cached values per lane
  __shared__ uint32_t tm_base_s;
  if (kTmem && wid == 0) tmem::alloc(&tm_base_s);

  // ---------------------------------------------------------------- phase 0
  // host_io: the per-structure arrays may live in page-locked HOST memory (zero-copy
  // over PCIe, `d3b200_batch_vjp_host_*`), so every global access of them is fully
  // coalesced: positions are staged through shared memory and the outputs are gathered
  // into the caller's atom order before they are written as contiguous blocks.
  const bool hio = A.host_io != 0;                   // CTA-uniform
  // raw positions: in the reduction tiles when they fit (idle until phase 1), else in acc
  const bool pos_in_red = (size_t)3 * N <= (size_t)nw * 2 * kRedV * kRedStride;
  const S* spos = hio ? (pos_in_red ? red : acc) : pos;
  // the first global loads of the CTA are issued before any set-up work, so that their latency (HBM, or PCIe
  // in the zero-copy form) overlaps it: this thread's first atomic number and, with resident inputs, its position
  const int64_t z_first = t < N ? numbers[t] : 0;
  S p_first[3] = {S(0), S(0), S(0)};
  if (!hio && t < N) { p_first[0] = pos[3 * t]; p_first[1] = pos[3 * t + 1]; p_first[2] = pos[3 * t + 2]; }
  for (int k = t; k < kMaxZ; k += nt) cnt[k] = 0;
  if (t == 0) misc[2] = 0;
  if (!hio || pos_in_red)
    for (int k = t; k < nw * 5 * N; k += nt) acc[k] = S(0);
  if (kTmem) tmem::fence_before_sync();
  __syncthreads();
  uint32_t tm_base = 0;
  if (kTmem) { tmem::fence_after_sync(); tm_base = tm_base_s; }
  int last_real = 0;
  for (int k = t; k < N; k += nt) {
    int Z = (int)(k == t ? z_first : numbers[k]);
    if (Z < 0 || Z >= kMaxZ) Z = 0;
    sorigZ[k] = (unsigned char)Z;
    if (Z > 0) {
      atomicAdd(&cnt[Z], 1);
      last_real = k + 1;
    } else if (!hio) {
      size_t o = (size_t)b * N + k;
      if (A.energy) A.energy[o] = S(0);
      if (WANT_G) { A.grad[3 * o] = S(0); A.grad[3 * o + 1] = S(0); A.grad[3 * o + 2] = S(0); }
    }
  }
  if (hio && last_real) atomicMax(&misc[2], last_real);
  __syncthreads();
  if (hio) {
    // host_io: only the positions up to the last real atom cross the bus (a padded batch of 50..120-atom
    // structures is 29 % padding), as one coalesced block
    S* dst = pos_in_red ? red : acc;
    const int m3 = 3 * misc[2];
    for (int m = t; m < m3; m += nt) dst[m] = pos[m];
  }
  if (t < 32) {
    // element histogram -> sorted offsets and group list, one warp, 4 elements per lane
    int c[4], present = 0, total = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int z = 4 * t + k;
      c[k] = (z >= 1 && z < kMaxZ) ? cnt[z] : 0;
      present += c[k] > 0;
      total += c[k];
    }
    int run = total, ng = present;          // inclusive scans over the lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int r2 = __shfl_up_sync(0xffffffffu, run, o), g2 = __shfl_up_sync(0xffffffffu, ng, o);
      if (t >= o) { run += r2; ng += g2; }
    }
    const int ntot = __shfl_sync(0xffffffffu, run, 31), gtot = __shfl_sync(0xffffffffu, ng, 31);
    run -= total;
    ng -= present;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int z = 4 * t + k;
      if (z >= 1 && z < kMaxZ) {
        zstart[z] = run;
        if (c[k] > 0) { gZ[ng] = z; gstart[ng] = run; zgrp[z] = ng; ++ng; run += c[k]; }
      }
    }
    if (t == 0) { gstart[gtot] = ntot; misc[0] = gtot; misc[1] = ntot; }
  }
  __syncthreads();
  const int ng = misc[0];
  const int n = misc[1];
  for (int k = t; k < N; k += nt) {
    const int Z = sorigZ[k];
    if (Z == 0) continue;
    int p = zstart[Z];
    for (int m = 0; m < k; ++m) p += (sorigZ[m] == Z);
    sperm[p] = (unsigned short)k;
    sZ[p] = (unsigned char)Z;
    sgrp[p] = (unsigned char)zgrp[Z];
    S rc = A.rcov_per_element ? A.rcov[Z] : A.rcov[(size_t)b * N + k];
    S q = A.r4r2_per_element ? A.r4r2[Z] : A.r4r2[(size_t)b * N + k];
    Atom4<S> a4;
    if (!hio && k == t) { a4.x = p_first[0]; a4.y = p_first[1]; a4.z = p_first[2]; }
    else { a4.x = spos[3 * k]; a4.y = spos[3 * k + 1]; a4.z = spos[3 * k + 2]; }
    a4.rk = A.kcn * rc;
    sA[p] = a4;
    Pair2<S> b2;
    b2.a = sqr(q);
    b2.b = (WANT_G && A.g) ? A.g[(size_t)b * N + k] : S(1);
    sB[p] = b2;
  }
  // reference systems per group (index of the last valid reference CN + 1); `cnt` is free from here on
  int* gnref = cnt;
  for (int gk = t; gk < ng; gk += nt) {
    int nr = 1;
#pragma unroll
    for (int a = 0; a < kNRef; ++a) nr = (A.refcn[gZ[gk] * kNRef + a] >= S(0)) ? a + 1 : nr;
    gnref[gk] = (D3_V2_NREF != 0) ? nr : kNRef;
  }
  __syncthreads();
  if (hio && !pos_in_red) {                          // (the staged positions are dead from here on)
    for (int k = t; k < nw * 5 * N; k += nt) acc[k] = S(0);
    __syncthreads();
  }

  const S tiny = tiny_of<S>();
  const S eps = eps_of<S>();
  const int nb = (n + 15) >> 4;
  S* accw = acc + (size_t)wid * 5 * N;
  S* redw = red + (size_t)wid * 2 * kRedV * kRedStride;

  // Each warp takes a contiguous share of the linearised upper triangle (row
  // block rb contributes its n-1-16rb partners j = 16rb+1 .. n-1, plus OVH
  // units of fixed per-block overhead so that warps with many short row
  // blocks -- more U contractions, more row epilogues -- get fewer partners).
#ifndef D3_V2_AL14
#define D3_V2_AL14 8
#endif
#define D3_PIECES_BEGIN(OVH) D3_PIECES_BEGIN_A(OVH, D3_V2_AL14)
#define D3_PIECES_BEGIN_A(OVH, AL)                                              \
  {                                                                             \
    const int T_ = nb * (n - 1 + (OVH)) - 8 * nb * (nb - 1);                    \
    const int wlo_ = min(T_, ((T_ * wid / nw) + (AL) - 1) & ~((AL) - 1));       \
    const int whi_ = wid == nw - 1 ? T_ : min(T_, ((T_ * (wid + 1) / nw) + (AL) - 1) & ~((AL) - 1)); \
    int pos_ = 0;                                                               \
    for (int rb = 0; rb < nb; ++rb) {                                           \
      const int i0 = rb << 4;                                                   \
      const int len_ = n - 1 - i0;                                              \
      const int a_ = max(wlo_, pos_ + (OVH)), b_ = min(whi_, pos_ + (OVH) + len_); \
      const int j0 = i0 + 1 + (a_ - pos_ - (OVH)), j1 = i0 + 1 + (b_ - pos_ - (OVH)); \
      pos_ += len_ + (OVH);                                                     \
      if (a_ >= b_) continue;                                                   \
      const int i = i0 + hl;                                                    \
      const bool irow = i < n;                                                  \
      const int ic = irow ? i : n - 1;
#define D3_PIECES_END }}

  // ---------------------------------------------------------------- phase 1
  // CN_i = sum_j [r_ij <= 25] / (1 + exp(-kcn (rc_ij / r_ij - 1))), pairs once
#ifndef D3_V2_CNB
#define D3_V2_CNB 4
#define D3_V2_PB 2
#define D3_V2_GB 3
#endif
  constexpr int CNB = D3_V2_CNB;
  constexpr int GB = 2;                     // partners per CN-backward step and half-warp (CNB % GB == 0)
  static_assert(CNB % GB == 0, "the two CN sweeps must walk the partners in the same order");
  int tslot = 0;                            // TMEM: values stored so far by this lane (same count in every lane)
  D3_PIECES_BEGIN(4)
  {
    const Atom4<S> ai = sA[ic];
    const unsigned jspan = (unsigned)max(j1 - i - 1, 0);   // partner j is valid  <=>  (unsigned)(j - i - 1) < jspan
    S cni = S(0);
    for (int jb0 = j0; jb0 < j1; jb0 += 2 * CNB) {  // warp-uniform trip count
      const int jb = jb0 + ph;
      S f[CNB];
      S dfc[CNB];
#pragma unroll
      for (int s = 0; s < CNB; ++s) {
        const int j = jb + 2 * s;
        const int jc = j < j1 ? j : j1 - 1;
        const Atom4<S> aj = sA[jc];
        S dx = ai.x - aj.x, dy = ai.y - aj.y, dz = ai.z - aj.z;
        S r2 = dx * dx + dy * dy + dz * dz;
        const bool pj = irow & ((unsigned)(j - i - 1) < jspan);
        const bool ok = pj & (r2 <= A.cn_cutoff2);
        S rinv = frsq(r2 + tiny);
        S rk = ai.rk + aj.rk;
        S e = fex(fmx(A.kcn - rk * rinv, S(-700)));
        S v = frcp(S(1) + e);
        f[s] = ok ? v : S(0);
        cni += f[s];
        if (kTmem) {
          // -(1/r) d f / d r = kcn rc r^-3 f (1 - f), f (1 - f) = e f^2; zero for pairs that do not count
          S df = (rk * (rinv * rinv * rinv)) * (v * (e * v));
          dfc[s] = (ok & (r2 >= eps)) ? df : S(0);
        }
      }
      if (kTmem) {
        if (tslot + CNB <= kTmemSlots) tmem::store<CNB>(tmem::addr(tm_base, wid, tslot * kWordsPerVal), dfc);
        tslot += CNB;
      }
      S tot = rows_reduce<CNB, S>(f, redw, lane);
      {
        const int r = red_row<CNB>(lane);
        const int j = jb0 + r / CNB + 2 * (r % CNB);
        if (red_owner<CNB>(lane) && j < j1) accw[j] += tot;
      }
    }
    cni += shx(cni, 16);
    __syncwarp();
    if (ph == 0 && irow) accw[i] += cni;
    __syncwarp();
  }
  D3_PIECES_END
  if (kTmem) tmem::wait_st();
  __syncthreads();

  // ---------------------------------------------------------------- phase 2
  for (int i = t; i < n; i += nt) {
    S cn = S(0);
    for (int w = 0; w < nw; ++w) { cn += acc[(size_t)w * 5 * N + i]; acc[(size_t)w * 5 * N + i] = S(0); }
    S wi[kNRef], dwi[kNRef];
    reference_weights<S>(A.refcn + sZ[i] * kNRef, cn, wi, dwi);
#pragma unroll
    for (int a = 0; a < kNRef; ++a) { sw[8 * i + a] = wi[a]; sdw[8 * i + a] = dwi[a]; }
    sw[8 * i + 7] = S(0); sdw[8 * i + 7] = S(0);
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase 3
  constexpr int PB = D3_V2_PB;              // partners per iteration and half-warp
  constexpr int PV = WANT_G ? 5 : 1;        // j-side values per partner
#ifndef D3_V2_OVH3
#define D3_V2_OVH3 18
#endif
#ifndef D3_V2_AL3
#define D3_V2_AL3 8
#endif
  D3_PIECES_BEGIN_A(D3_V2_OVH3, D3_V2_AL3)
  {
    const Atom4<S> ai = sA[ic];
    const Pair2<S> bi = sB[ic];
    const int Zi = sZ[ic];
    const S gi = bi.b;
    const S s3qi = sqr(S(3)) * bi.a;        // sqrt(3 q_i)
    S ei = S(0), gxi = S(0), gyi = S(0), gzi = S(0), di = S(0);
    const unsigned jspan = (unsigned)max(j1 - i - 1, 0);
    // rows of a block are sorted by element: the largest reference count among them bounds the a-loop
    const int nri = __reduce_max_sync(0xffffffffu, gnref[sgrp[ic]]);
    for (int grp = sgrp[j0]; grp < ng && gstart[grp] < j1; ++grp) {
      const int jlo = max(gstart[grp], j0), jhi = min(gstart[grp + 1], j1);
      const S* __restrict__ K = A.refc6 + ((size_t)Zi * kMaxZ + gZ[grp]) * (kNRef * kNRef);
      // everything below is instantiated for NR = 2, 3, 5, 7 columns (reference systems of the partner element)
      auto sweep = [&](auto nr_tag) {
        constexpr int NR = decltype(nr_tag)::value;
        constexpr int NP = (NR + 1) / 2;        // 128-bit loads per weight vector
        // U[b] = sum_a w_ia K[Zi,Zg,a,b],  Ud[b] = sum_a dw_ia K[Zi,Zg,a,b]:
        // half 0 contracts the weights, half 1 their CN derivatives, then swap
        S U[NR], Ud[NR];
        {
          const S* src = (WANT_G && ph == 1) ? sdw + 8 * ic : sw + 8 * ic;
          S mine[NR];
#pragma unroll
          for (int bb = 0; bb < NR; ++bb) mine[bb] = S(0);
#pragma unroll
          for (int a = 0; a < kNRef; ++a) {
            if (a < nri) {                      // warp-uniform
              const S wa = src[a];
#pragma unroll
              for (int bb = 0; bb < NR; ++bb) mine[bb] += wa * __ldg(K + a * kNRef + bb);
            }
          }
#pragma unroll
          for (int bb = 0; bb < NR; ++bb) {
            if (WANT_G) {
              S other = shx(mine[bb], 16);
              U[bb] = ph == 0 ? mine[bb] : other;
              Ud[bb] = ph == 0 ? other : mine[bb];
            } else {
              U[bb] = mine[bb];
            }
          }
        }
        S r06g = S(0), r08g = S(0), s8qg = S(0);
        if (QE) {
          const S tq = s3qi * sB[gstart[grp]].a;     // sqrt(3 q_i q_g)
          const S qq = tq * tq;
          const S r0 = A.a1 * tq + A.a2;
          const S r02 = r0 * r0;
          r06g = r02 * r02 * r02;
          r08g = r06g * r02;
          s8qg = A.s8 * qq;
        }
        for (int jb0 = jlo; jb0 < jhi; jb0 += 2 * PB) {
          S vals[PV * PB];
#pragma unroll
          for (int s = 0; s < PB; ++s) {
            const int j = jb0 + ph + 2 * s;
            const bool jin = j < jhi;
            const int jc = jin ? j : jhi - 1;
            const Atom4<S> aj = sA[jc];
            const Pair2<S> bj = sB[jc];
            S dx = ai.x - aj.x, dy = ai.y - aj.y, dz = ai.z - aj.z;
            S r2 = dx * dx + dy * dy + dz * dz;
            const bool pj = irow & jin & ((unsigned)(j - i - 1) < jspan);
            const bool ok = pj & (r2 <= A.cutoff2);
            const Pair2<S>* wj = reinterpret_cast<const Pair2<S>*>(sw + 8 * jc);
            S wv[2 * NP];
#pragma unroll
            for (int k = 0; k < NP; ++k) { Pair2<S> p = wj[k]; wv[2 * k] = p.a; wv[2 * k + 1] = p.b; }
            S c6 = U[0] * wv[0];
#pragma unroll
            for (int a = 1; a < NR; ++a) c6 += U[a] * wv[a];
            S r06, r08, s8q;
            if (QE) {
              r06 = r06g; r08 = r08g; s8q = s8qg;
            } else {
              S tq = s3qi * bj.a;             // sqrt(3 q_i q_j)
              S qq = tq * tq;
              S r0 = A.a1 * tq + A.a2;        // R0 = a1 sqrt(3 q_i q_j) + a2
              S r02 = r0 * r0;
              r06 = r02 * r02 * r02;
              r08 = r06 * r02;
              s8q = A.s8 * qq;
            }
            S r4 = r2 * r2;
            S r6 = r4 * r2;
            S r8 = r4 * r4;
            S d6 = r6 + r06, d8 = r8 + r08;
            S inv = frcp(d6 * d8);
            S t6 = inv * d8, t8 = inv * d6;
            S P = A.s6 * t6 + s8q * t8;
            S ep = ok ? c6 * P : S(0);
            vals[PV * s] = ep;
            ei += ep;
            if (WANT_G) {
              const Pair2<S>* dwj = reinterpret_cast<const Pair2<S>*>(sdw + 8 * jc);
              S dv[2 * NP];
#pragma unroll
              for (int k = 0; k < NP; ++k) { Pair2<S> p = dwj[k]; dv[2 * k] = p.a; dv[2 * k + 1] = p.b; }
              S dci = Ud[0] * wv[0], dcj = U[0] * dv[0];
#pragma unroll
              for (int a = 1; a < NR; ++a) { dci += Ud[a] * wv[a]; dcj += U[a] * dv[a]; }
              S gam = ok ? S(0.5) * (gi + bj.b) : S(0);
              S gP = gam * P;
              di -= gP * dci;
              S t62 = t6 * t6, t82 = t8 * t8;
              S dP = S(3) * (A.s6 * (r4 * t62)) + S(4) * (s8q * (r6 * t82));
              S f = (S(2) * gam) * (c6 * dP);
              S fx = f * dx, fy = f * dy, fz = f * dz;
              gxi += fx; gyi += fy; gzi += fz;
              vals[PV * s + 1 % PV] = -fx; vals[PV * s + 2 % PV] = -fy; vals[PV * s + 3 % PV] = -fz;
              vals[PV * s + 4 % PV] = -(gP * dcj);
            }
          }
          S tot = rows_reduce<PV * PB, S>(vals, redw, lane);
          {
            constexpr int V = PV * PB;
            const int r = red_row<V>(lane);
            const int v = r % V;
            const int j = jb0 + r / V + 2 * (v / PV);
            if (red_owner<V>(lane) && j < jhi) accw[(size_t)(v % PV) * N + j] += tot;
          }
        }
      };
      const int nrg = gnref[grp];
#ifndef D3_V2_NR3
#define D3_V2_NR3 1
#endif
      if (nrg <= 2) sweep(std::integral_constant<int, 2>{});
      else if (D3_V2_NR3 && nrg == 3) sweep(std::integral_constant<int, 3>{});
      else if (nrg <= 5) sweep(std::integral_constant<int, 5>{});
      else sweep(std::integral_constant<int, kNRef>{});
    }
    ei += shx(ei, 16);
    if (WANT_G) { gxi += shx(gxi, 16); gyi += shx(gyi, 16); gzi += shx(gzi, 16); di += shx(di, 16); }
    __syncwarp();
    if (ph == 0 && irow) {
      accw[i] += ei;
      if (WANT_G) { accw[N + i] += gxi; accw[2 * N + i] += gyi; accw[3 * N + i] += gzi; accw[4 * N + i] += di; }
    }
    __syncwarp();
  }
  D3_PIECES_END
  __syncthreads();
  S* se = sdecn + N;                                // second half of sB: energies in the caller's order
  for (int i = t; i < n; i += nt) {
    S e = S(0), d = S(0);
    for (int w = 0; w < nw; ++w) { e += acc[(size_t)w * 5 * N + i]; d += acc[((size_t)w * 5 + 4) * N + i]; }
    if (hio) se[sperm[i]] = S(-0.5) * e;
    else if (A.energy) A.energy[(size_t)b * N + sperm[i]] = S(-0.5) * e;
    sdecn[i] = d;
  }
  __syncthreads();
  if (hio && A.energy)
    for (int k = t; k < N; k += nt) A.energy[(size_t)b * N + k] = sorigZ[k] ? se[k] : S(0);
  if (!WANT_G) return;

  // ---------------------------------------------------------------- phase 4
  // CN chain: (dL/dCN_i + dL/dCN_j) df_ij/dx, df/dx_i = -kcn rc r^-3 f(1-f) (x_i - x_j)
  tslot = 0;
  D3_PIECES_BEGIN(4)
  {
    const Atom4<S> ai = sA[ic];
    const S dci = sdecn[ic];
    const unsigned jspan = (unsigned)max(j1 - i - 1, 0);
    S gxi = S(0), gyi = S(0), gzi = S(0);
    for (int jb0 = j0; jb0 < j1; jb0 += 2 * CNB) {  // warp-uniform trip count, same steps as the CN sweep
      S dfc[CNB];
      const bool cached = kTmem && (tslot + CNB <= kTmemSlots);   // warp-uniform
      if (cached) tmem::load<CNB>(tmem::addr(tm_base, wid, tslot * kWordsPerVal), dfc);
      tslot += CNB;
#pragma unroll
      for (int h = 0; h < CNB / GB; ++h) {
        const int jh0 = jb0 + 2 * GB * h;              // partners jh0 + ph + 2 s, s < GB
        S vals[3 * GB];
#pragma unroll
        for (int s = 0; s < GB; ++s) {
          const int j = jh0 + ph + 2 * s;
          const int jc = j < j1 ? j : j1 - 1;
          const Atom4<S> aj = sA[jc];
          S dx = ai.x - aj.x, dy = ai.y - aj.y, dz = ai.z - aj.z;
          S df;
          if (cached) {
            df = dfc[GB * h + s];
          } else {
            // structure too large for the TMEM slots of this lane: recompute (kept out of line so that the
            // common loop body stays small)
            const bool pj = irow & ((unsigned)(j - i - 1) < jspan);
            df = cn_df_recompute<S>(dx, dy, dz, ai.rk + aj.rk, pj, A.cn_cutoff2, A.kcn);
          }
          S c = (dci + sdecn[jc]) * df;                 // = -(dL/dCN_i + dL/dCN_j) f'(r) / r
          S cx = c * dx, cy = c * dy, cz = c * dz;
          gxi -= cx; gyi -= cy; gzi -= cz;
          vals[3 * s] = cx; vals[3 * s + 1] = cy; vals[3 * s + 2] = cz;
        }
        if (jh0 < j1) {                                  // warp-uniform: the second half of the last step may be empty
          S tot = rows_reduce<3 * GB, S>(vals, redw, lane);
          constexpr int V = 3 * GB;
          const int r = red_row<V>(lane);
          const int v = r % V;
          const int j = jh0 + r / V + 2 * (v / 3);
          if (red_owner<V>(lane) && j < j1) accw[(size_t)(1 + v % 3) * N + j] += tot;
        }
      }
    }
    gxi += shx(gxi, 16); gyi += shx(gyi, 16); gzi += shx(gzi, 16);
    __syncwarp();
    if (ph == 0 && irow) { accw[N + i] += gxi; accw[2 * N + i] += gyi; accw[3 * N + i] += gzi; }
    __syncwarp();
  }
  D3_PIECES_END
  __syncthreads();
  if (kTmem && wid == 0) tmem::dealloc(tm_base);      // every warp has read its last value
#undef D3_PIECES_BEGIN
#undef D3_PIECES_BEGIN_A
#undef D3_PIECES_END
  // host_io: gradient gathered into the caller's order in the (spent) energy slots of
  // warps 0..2, then written as one contiguous [N][3] block, zeros on padding
  static_assert(kV2Warps >= 3, "three spare accumulator rows needed");
  S* gxo = acc;
  S* gyo = acc + (size_t)5 * N;
  S* gzo = acc + (size_t)10 * N;
  for (int i = t; i < n; i += nt) {
    S gx = S(0), gy = S(0), gz = S(0);
    for (int w = 0; w < nw; ++w) {
      const S* a = acc + (size_t)w * 5 * N;
      gx += a[N + i]; gy += a[2 * N + i]; gz += a[3 * N + i];
    }
    const int k = sperm[i];
    if (hio) {
      gxo[k] = gx; gyo[k] = gy; gzo[k] = gz;
    } else {
      const size_t o = (size_t)b * N + k;
      A.grad[3 * o] = gx; A.grad[3 * o + 1] = gy; A.grad[3 * o + 2] = gz;
    }
  }
  if (!hio) return;
  __syncthreads();
  S* gout = A.grad + (size_t)b * N * 3;
  for (int m = t; m < 3 * N; m += nt) {
    const int k = m / 3, c = m - 3 * k;
    const S v = c == 0 ? gxo[k] : (c == 1 ? gyo[k] : gzo[k]);
    gout[m] = sorigZ[k] ? v : S(0);
  }
}

}  // namespace d3
