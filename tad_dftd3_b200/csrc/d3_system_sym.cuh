// Pairs-once ("symmetric") variants of the tiled sweeps.
//
// The full-row sweeps of d3_system.cuh visit every ordered pair, i.e. they
// evaluate each pair twice.  Here a CTA owns a row tile I and walks only the
// tiles J >= I (j > i inside the diagonal tile); each pair is evaluated once,
// the i-side sums stay in the registers of the row's thread, and the j-side
// values of 32 rows x 16 partners are staged in a per-warp shared tile and
// summed per partner by the lanes.  Both sides are added to the global arrays
// with fp64 atomics (one per row and quantity per CTA, one per partner and
// quantity per 32 x 16 chunk), so the arrays must be zero before a sweep, and
// on several GPUs they are summed with all-reduce.  Row tiles are dealt to
// ranks round-robin (`tile_stride`), which balances the triangle.
#pragma once
#include "d3_atm.cuh"

namespace d3 {

#ifndef D3_SYM_CHUNK
#define D3_SYM_CHUNK 8
#endif
// 1: the pair sweep reads the pre-contracted vectors T_j[e_i], Td_j[e_i] (64 B each, one partner per warp step, at most
// `nelem` distinct rows per load) straight from global memory through L1 instead of staging 128 x nelem x 128 B per
// j-tile in shared memory: 66 -> 34 kB per CTA, i.e. 4 instead of 3 CTAs per SM, and 32 kB less to copy per tile.
#ifndef D3_SYM_T_GLOBAL
#define D3_SYM_T_GLOBAL 1
#endif
constexpr int kSymChunk = D3_SYM_CHUNK;  // partners per transposed chunk
constexpr int kSymParts = 32 / kSymChunk; // lanes sharing one partner column
constexpr int kSymPitch = 33;            // 32 rows + 1: conflict-free column reads
// (measured on the 50k-atom cluster: 1 tile per barrier pair 2.39 ms, 2 tiles 2.45, 4 tiles 2.45 -- the warps'
// imbalance does not average out over a batch, so the default stays at one tile)
#ifndef D3_SYM_BATCH
#define D3_SYM_BATCH 1
#endif
constexpr int kSymBatch = D3_SYM_T_GLOBAL != 0 ? D3_SYM_BATCH : 1;   // j-tiles staged per pair of barriers (pair sweep)

template <typename S>
struct SymArgs {
  SysArgs<S> sys;
  int tile_stride;          // this launch owns row tiles row_begin/128 + k * tile_stride
};

template <typename S>
__device__ __forceinline__ bool sym_row_tile(const SymArgs<S>& P, int& itile) {
  // grid = (jsplit, row tiles): the CTAs of the early row tiles (which walk the most j-tiles of the
  // triangle J >= I) are scheduled first, the cheap ones make up the tail
  itile = P.sys.row_begin / kSysTile + blockIdx.y * P.tile_stride;
  return itile * kSysTile < P.sys.row_end;
}

// squared distance of a point to a box {min xyz, max xyz}
template <typename S>
__device__ __forceinline__ S box_point_dist2(const S* box, const Atom4<S>& a) {
  S d2 = S(0);
  const S x[3] = {a.x, a.y, a.z};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const S lo = box[k] - x[k], hi = x[k] - box[3 + k];
    S d = lo > hi ? lo : hi;
    d = d > S(0) ? d : S(0);
    d2 += d * d;
  }
  return d2;
}

// Candidate partners of one warp in the staged j-tile: those within `cut2` of the box of the
// warp's 32 rows (a 32-atom box test alone lets through about as many partners that no row can
// reach as ones that count).  The warp compacts their tile-local indices, in order, into
// `cand` (padded with 255 up to a multiple of the chunk) and returns their number.  On the
// diagonal tile only partners at or behind the warp's first row can have j > i.
template <typename S>
__device__ __forceinline__ int sym_candidates(const Atom4<S>* sj, const S* box, const S* wbox, int jt, int itile,
                                              int wsub, S cut2, unsigned char* cand, int lane) {
  int n = 0;
#pragma unroll
  for (int sub = 0; sub < kSysTile / kSysSub; ++sub) {
    if (jt == itile && sub < wsub) continue;
    if (box_dist2(wbox, box + ((size_t)jt * 4 + sub) * 6) > cut2) continue;   // warp-uniform
    const int jl = sub * kSysSub + lane;
    const bool ok = box_point_dist2(wbox, sj[jl]) <= cut2;
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (ok) cand[n + __popc(m & ((1u << lane) - 1))] = (unsigned char)jl;
    n += __popc(m);
  }
  if (lane < kSymChunk) cand[n + lane] = 255;        // tail of the last chunk
  __syncwarp();
  return n;
}

// The j-tiles a CTA walks are itile + blockIdx.x + k * jsplit, k = 0, 1, ...; most of them are
// out of reach.  Lanes test 32 of them at once (tile box against the row tile's box) and the
// warp then visits only the set bits -- every warp of the CTA computes the same mask, so the
// CTA-wide staging barriers stay aligned.
template <typename S>
__device__ __forceinline__ unsigned sym_tile_mask(const S* box, const S* ibox, int jt_first, int jsplit, int k0,
                                                  int ntiles, S cut2, int lane) {
  const int jt = jt_first + (k0 + lane) * jsplit;
  bool ok = false;
  if (jt < ntiles) {
    S jbox[6];
    tile_box(box, jt, jbox);
    ok = box_dist2(ibox, jbox) <= cut2;
  }
  return __ballot_sync(0xffffffffu, ok);
}

// ---------------------------------------------------------------- CN sweep
template <typename S>
__global__ void __launch_bounds__(kSysTile) sym_cn_kernel(SymArgs<S> P) {
  const SysArgs<S>& A = P.sys;
  __shared__ Atom4<S> sj[kSysTile];
  __shared__ S sS[kSysTile / 32][kSymChunk * kSymPitch];
  __shared__ unsigned char scand[kSysTile / 32][kSysTile + kSymChunk];
  int itile;
  if (!sym_row_tile(P, itile)) return;
  const int t = threadIdx.x, lane = t & 31, wsub = t >> 5;
  unsigned char* cand = scand[wsub];
  const int i = itile * kSysTile + t;
  const Atom4<S> ai = A.atoms[i];
  const S tiny = tiny_of<S>();
  S ibox[6], wbox[6];
  tile_box(A.box, itile, ibox);
#pragma unroll
  for (int k = 0; k < 6; ++k) wbox[k] = A.box[((size_t)itile * 4 + wsub) * 6 + k];
  const int ntiles = A.natoms / kSysTile;
  S* S_ = sS[wsub];
  S cn = S(0);
  const int jt_first = itile + blockIdx.x;
  for (int k0 = 0; jt_first + k0 * A.jsplit < ntiles; k0 += 32)
  for (unsigned tmask = sym_tile_mask<S>(A.box, ibox, jt_first, A.jsplit, k0, ntiles, A.cn_cutoff2, lane); tmask;
       tmask &= tmask - 1) {
    const int jt = jt_first + (k0 + __ffs(tmask) - 1) * A.jsplit;      // CTA-uniform
    __syncthreads();
    sj[t] = A.atoms[(size_t)jt * kSysTile + t];
    __syncthreads();
    {
      const int ncand = sym_candidates<S>(sj, A.box, wbox, jt, itile, wsub, A.cn_cutoff2, cand, lane);
      for (int c0 = 0; c0 < ncand; c0 += kSymChunk) {
#pragma unroll 4
        for (int jj = 0; jj < kSymChunk; ++jj) {
          const int jl = cand[c0 + jj];
          S v = S(0);
          if (jl < kSysTile) {                            // warp-uniform (255 = chunk padding)
            const int j = jt * kSysTile + jl;
            const Atom4<S> aj = sj[jl];
            S dx = ai.x - aj.x, dy = ai.y - aj.y, dz = ai.z - aj.z;
            S r2 = dx * dx + dy * dy + dz * dz;
            if (r2 <= A.cn_cutoff2 && j > i) {
              S rinv = frsq(r2 + tiny);
              S x = fmx(A.kcn - (ai.rk + aj.rk) * rinv, S(-700));
              v = frcp(S(1) + fex(x));
              cn += v;
            }
          }
          S_[jj * kSymPitch + lane] = v;
        }
        __syncwarp();
        {
          // kSymParts lanes share partner column `col`, each sums kSymChunk of the 32 rows
          const int col = lane % kSymChunk, r0 = (lane / kSymChunk) * kSymChunk;
          S s = S(0);
#pragma unroll
          for (int r = 0; r < kSymChunk; ++r) s += S_[col * kSymPitch + r0 + r];
#pragma unroll
          for (int o = kSymChunk; o < 32; o <<= 1) s += shx(s, o);
          if (lane < kSymChunk && s != S(0)) red_add(A.cn + (size_t)jt * kSysTile + cand[c0 + col], s);
        }
        __syncwarp();
      }
    }
  }
  if (cn != S(0)) red_add(A.cn + i, cn);
}

// ---------------------------------------------------------------- pair sweep (T vectors)
template <typename S, bool WANT_G>
__global__ void __launch_bounds__(kSysTile, 4) sym_pair_t_kernel(SymArgs<S> P) {
  const SysArgs<S>& A = P.sys;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int E = A.nelem;
  constexpr bool kTG = D3_SYM_T_GLOBAL != 0;
  constexpr int NB = kSymBatch;
  __shared__ int sjts[kSymBatch];                                // j-tiles staged per pair of CTA barriers
  Atom4<S>* sj = reinterpret_cast<Atom4<S>*>(smem_raw);        // [NB][128]
  Pair2<S>* sb = reinterpret_cast<Pair2<S>*>(sj + NB * kSysTile);
  S* sT = reinterpret_cast<S*>(sb + NB * kSysTile);                 // [128][E][8]  T vectors of the j tile (staged form)
  S* sTd = sT + (kTG ? 0 : (size_t)kSysTile * E * 8);          // [128][E][8]  Td vectors (dw-contracted)
  S* sSall = sTd + (kTG ? 0 : (size_t)kSysTile * E * 8);       // [4 warps][3][chunk][33]
  S* sXi = sSall + (size_t)4 * 3 * kSymChunk * kSymPitch;      // [4 warps][32][4] rows' x,y,z
  int* sej = reinterpret_cast<int*>(sXi + 4 * 32 * 4);
  int itile;
  if (!sym_row_tile(P, itile)) return;
  const int t = threadIdx.x, lane = t & 31, wsub = t >> 5;
  const int i = itile * kSysTile + t;
  const Atom4<S> ai = A.atoms[i];
  const Pair2<S> bi = A.aux[i];
  const int ei = A.eidx[i];
  const int eic = ei < 0 ? 0 : ei;
  const S tiny = tiny_of<S>();
  S wi[kNRef], dwi[kNRef];
#pragma unroll
  for (int a = 0; a < kNRef; ++a) { wi[a] = A.w[8 * (size_t)i + a]; dwi[a] = A.dw[8 * (size_t)i + a]; }
  const S* tdvec = A.tvec + (size_t)A.natoms * E * 8;
  S ibox[6], wbox[6];
  tile_box(A.box, itile, ibox);
#pragma unroll
  for (int k = 0; k < 6; ++k) wbox[k] = A.box[((size_t)itile * 4 + wsub) * 6 + k];
  S* S_ = sSall + (size_t)wsub * 3 * kSymChunk * kSymPitch;
  S* Xi = sXi + wsub * 32 * 4;
  unsigned char* cand = reinterpret_cast<unsigned char*>(sej + NB * kSysTile) + wsub * (kSysTile + kSymChunk);
  Xi[lane * 4] = ai.x; Xi[lane * 4 + 1] = ai.y; Xi[lane * 4 + 2] = ai.z;
  __syncwarp();
  const int ntiles = A.natoms / kSysTile;
  const S gi = bi.b;
  const S s3qi = sqr(S(3)) * bi.a;
  S e_i = S(0), gx = S(0), gy = S(0), gz = S(0), decn = S(0);
  const int jt_first = itile + blockIdx.x;
  for (int k0 = 0; jt_first + k0 * A.jsplit < ntiles; k0 += 32)
  for (unsigned tmask = sym_tile_mask<S>(A.box, ibox, jt_first, A.jsplit, k0, ntiles, A.cutoff2, lane); tmask;) {
    // up to NB reachable j-tiles are staged between one pair of CTA barriers: with the T vectors read from
    // global memory a tile is 6.5 kB, and the warps' different candidate counts average out over the batch
    int nbt = 0;
    __syncthreads();
#pragma unroll
    for (int bb = 0; bb < NB; ++bb) {
      if (!tmask) break;
      const int jtb = jt_first + (k0 + __ffs(tmask) - 1) * A.jsplit;   // CTA-uniform
      tmask &= tmask - 1;
      ++nbt;
      if (t == 0) sjts[bb] = jtb;
      const size_t j = (size_t)jtb * kSysTile + t;
      sj[bb * kSysTile + t] = A.atoms[j];
      sb[bb * kSysTile + t] = A.aux[j];
      sej[bb * kSysTile + t] = A.eidx[j];
      if (!kTG) {
        const Pair2<S>* src = reinterpret_cast<const Pair2<S>*>(A.tvec + j * E * 8);
        Pair2<S>* dst = reinterpret_cast<Pair2<S>*>(sT + (size_t)t * E * 8);
        for (int k = 0; k < 4 * E; ++k) dst[k] = src[k];
        if (WANT_G) {
          const Pair2<S>* srcd = reinterpret_cast<const Pair2<S>*>(tdvec + j * E * 8);
          Pair2<S>* dstd = reinterpret_cast<Pair2<S>*>(sTd + (size_t)t * E * 8);
          for (int k = 0; k < 4 * E; ++k) dstd[k] = srcd[k];
        }
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int bb = 0; bb < nbt; ++bb) {
      const int jt = sjts[bb];
      const Atom4<S>* sjb = sj + bb * kSysTile;
      const Pair2<S>* sbb = sb + bb * kSysTile;
      const int* sejb = sej + bb * kSysTile;
      const int ncand = sym_candidates<S>(sjb, A.box, wbox, jt, itile, wsub, A.cutoff2, cand, lane);
      for (int c0 = 0; c0 < ncand; c0 += kSymChunk) {
#pragma unroll 2
        for (int jj = 0; jj < kSymChunk; ++jj) {
          const int jlr = cand[c0 + jj];                  // 255 = chunk padding (warp-uniform)
          const int jl = jlr & (kSysTile - 1);
          const int j = jt * kSysTile + jl;
          const int ej = jlr < kSysTile ? sejb[jl] : -1;
          const Atom4<S> aj = sjb[jl];
          S dx = ai.x - aj.x, dy = ai.y - aj.y, dz = ai.z - aj.z;
          S r2 = dx * dx + dy * dy + dz * dz;
          S ep = S(0), fj = S(0), dj = S(0);
          if (ej >= 0 && ei >= 0 && r2 <= A.cutoff2 && j > i) {
            r2 = r2 + tiny;
            const Pair2<S> bj = sbb[jl];
            const Pair2<S>* tj = kTG ? reinterpret_cast<const Pair2<S>*>(A.tvec + ((size_t)j * E + eic) * 8)
                                     : reinterpret_cast<const Pair2<S>*>(sT + ((size_t)jl * E + eic) * 8);
            S tv[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) { Pair2<S> p = tj[k]; tv[2 * k] = p.a; tv[2 * k + 1] = p.b; }
            S c6 = wi[0] * tv[0];
#pragma unroll
            for (int a = 1; a < kNRef; ++a) c6 += wi[a] * tv[a];
            S tq = s3qi * bj.a;
            S qq = tq * tq;
            S r0 = A.a1 * tq + A.a2;
            S r02 = r0 * r0;
            S r06 = r02 * r02 * r02;
            S r08 = r06 * r02;
            S r4 = r2 * r2;
            S r6 = r4 * r2;
            S r8 = r4 * r4;
            S d6 = r6 + r06, d8 = r8 + r08;
            S inv = frcp(d6 * d8);
            S t6 = inv * d8, t8 = inv * d6;
            S s8q = A.s8 * qq;
            S Pp = A.s6 * t6 + s8q * t8;
            ep = c6 * Pp;
            e_i += ep;
            if (WANT_G) {
              S dc6i = dwi[0] * tv[0];
#pragma unroll
              for (int a = 1; a < kNRef; ++a) dc6i += dwi[a] * tv[a];
              // dC6_ij/dCN_j = w_i . Td_j[e_i]   (Td_j[e][a] = sum_b K[Z_e,Z_j,a,b] dw_jb)
              const Pair2<S>* tdj = kTG ? reinterpret_cast<const Pair2<S>*>(tdvec + ((size_t)j * E + eic) * 8)
                                        : reinterpret_cast<const Pair2<S>*>(sTd + ((size_t)jl * E + eic) * 8);
              S td[8];
#pragma unroll
              for (int k = 0; k < 4; ++k) { Pair2<S> p = tdj[k]; td[2 * k] = p.a; td[2 * k + 1] = p.b; }
              S dc6j = wi[0] * td[0];
#pragma unroll
              for (int a = 1; a < kNRef; ++a) dc6j += wi[a] * td[a];
              S gam = S(0.5) * (gi + bj.b);
              S gP = gam * Pp;
              decn -= gP * dc6i;
              dj = -(gP * dc6j);
              S t62 = t6 * t6, t82 = t8 * t8;
              S dP = S(3) * (A.s6 * (r4 * t62)) + S(4) * (s8q * (r6 * t82));
              fj = (S(2) * gam) * (c6 * dP);
              gx += fj * dx; gy += fj * dy; gz += fj * dz;
            }
          }
          S_[jj * kSymPitch + lane] = ep;
          if (WANT_G) {
            S_[(kSymChunk + jj) * kSymPitch + lane] = fj;
            S_[(2 * kSymChunk + jj) * kSymPitch + lane] = dj;
          }
        }
        __syncwarp();
        {
          const int col = lane % kSymChunk, r0 = (lane / kSymChunk) * kSymChunk;
          const int jl = cand[c0 + col] & (kSysTile - 1);  // (padding columns sum to zero: no atomics)
          const Atom4<S> aj = sjb[jl];
          S se = S(0), sf = S(0), sfx = S(0), sfy = S(0), sfz = S(0), sd = S(0);
#pragma unroll 4
          for (int r = 0; r < kSymChunk; ++r) {
            const int row = r0 + r;
            se += S_[col * kSymPitch + row];
            if (WANT_G) {
              const S f = S_[(kSymChunk + col) * kSymPitch + row];
              sf += f;
              sfx += f * Xi[row * 4]; sfy += f * Xi[row * 4 + 1]; sfz += f * Xi[row * 4 + 2];
              sd += S_[(2 * kSymChunk + col) * kSymPitch + row];
            }
          }
#pragma unroll
          for (int o = kSymChunk; o < 32; o <<= 1) {
            se += shx(se, o);
            if (WANT_G) { sf += shx(sf, o); sfx += shx(sfx, o); sfy += shx(sfy, o); sfz += shx(sfz, o); sd += shx(sd, o); }
          }
          if (lane < kSymChunk && se != S(0)) {
            const size_t j = (size_t)jt * kSysTile + jl;
            red_add(A.energy + j, S(-0.5) * se);
            if (WANT_G) {
              // grad_j = -sum_i f_ij (x_i - x_j)
              red_add(A.grad + 3 * j, sf * aj.x - sfx);
              red_add(A.grad + 3 * j + 1, sf * aj.y - sfy);
              red_add(A.grad + 3 * j + 2, sf * aj.z - sfz);
              red_add(A.decn + j, sd);
            }
          }
        }
        __syncwarp();
      }
    }
  }
  if (e_i != S(0)) {
    red_add(A.energy + i, S(-0.5) * e_i);
    if (WANT_G) {
      red_add(A.grad + 3 * (size_t)i, gx); red_add(A.grad + 3 * (size_t)i + 1, gy);
      red_add(A.grad + 3 * (size_t)i + 2, gz); red_add(A.decn + i, decn);
    }
  }
}

template <typename S>
inline size_t sym_pair_smem_bytes(int nelem) {
  const size_t tstage = D3_SYM_T_GLOBAL != 0 ? 0 : (size_t)kSysTile * nelem * 16;
  return kSymBatch * kSysTile * (sizeof(Atom4<S>) + sizeof(Pair2<S>) + sizeof(int)) +
         (tstage + 4 * 3 * kSymChunk * kSymPitch + 4 * 32 * 4) * sizeof(S) +
         4 * (kSysTile + kSymChunk) + 64;
}

// ---------------------------------------------------------------- CN backward
template <typename S>
__global__ void __launch_bounds__(kSysTile) sym_cnbwd_kernel(SymArgs<S> P) {
  const SysArgs<S>& A = P.sys;
  __shared__ Atom4<S> sj[kSysTile];
  __shared__ S sd[kSysTile];
  __shared__ S sS[kSysTile / 32][kSymChunk * kSymPitch];
  __shared__ S sXi[kSysTile / 32][32 * 4];
  __shared__ unsigned char scand[kSysTile / 32][kSysTile + kSymChunk];
  int itile;
  if (!sym_row_tile(P, itile)) return;
  const int t = threadIdx.x, lane = t & 31, wsub = t >> 5;
  unsigned char* cand = scand[wsub];
  const int i = itile * kSysTile + t;
  const Atom4<S> ai = A.atoms[i];
  const S dci = A.decn[i];
  const S tiny = tiny_of<S>(), eps = eps_of<S>();
  S ibox[6], wbox[6];
  tile_box(A.box, itile, ibox);
#pragma unroll
  for (int k = 0; k < 6; ++k) wbox[k] = A.box[((size_t)itile * 4 + wsub) * 6 + k];
  S* S_ = sS[wsub];
  S* Xi = sXi[wsub];
  Xi[lane * 4] = ai.x; Xi[lane * 4 + 1] = ai.y; Xi[lane * 4 + 2] = ai.z;
  __syncwarp();
  const int ntiles = A.natoms / kSysTile;
  S gx = S(0), gy = S(0), gz = S(0);
  const int jt_first = itile + blockIdx.x;
  for (int k0 = 0; jt_first + k0 * A.jsplit < ntiles; k0 += 32)
  for (unsigned tmask = sym_tile_mask<S>(A.box, ibox, jt_first, A.jsplit, k0, ntiles, A.cn_cutoff2, lane); tmask;
       tmask &= tmask - 1) {
    const int jt = jt_first + (k0 + __ffs(tmask) - 1) * A.jsplit;      // CTA-uniform
    __syncthreads();
    sj[t] = A.atoms[(size_t)jt * kSysTile + t];
    sd[t] = A.decn[(size_t)jt * kSysTile + t];
    __syncthreads();
    {
      const int ncand = sym_candidates<S>(sj, A.box, wbox, jt, itile, wsub, A.cn_cutoff2, cand, lane);
      for (int c0 = 0; c0 < ncand; c0 += kSymChunk) {
#pragma unroll 2
        for (int jj = 0; jj < kSymChunk; ++jj) {
          const int jlr = cand[c0 + jj];                  // 255 = chunk padding (warp-uniform)
          const int jl = jlr & (kSysTile - 1);
          const int j = jt * kSysTile + jl;
          const Atom4<S> aj = sj[jl];
          S dx = ai.x - aj.x, dy = ai.y - aj.y, dz = ai.z - aj.z;
          S r2 = dx * dx + dy * dy + dz * dz;
          S c = S(0);
          if (jlr < kSysTile && r2 <= A.cn_cutoff2 && r2 >= eps && j > i) {
            S rinv = frsq(r2 + tiny);
            S rk = ai.rk + aj.rk;
            S e = fex(fmx(A.kcn - rk * rinv, S(-700)));
            S f = frcp(S(1) + e);
            S df = (rk * (rinv * rinv * rinv)) * (f * (e * f));
            c = (dci + sd[jl]) * df;               // grad_i -= c d, grad_j += c d
            gx -= c * dx; gy -= c * dy; gz -= c * dz;
          }
          S_[jj * kSymPitch + lane] = c;
        }
        __syncwarp();
        {
          const int col = lane % kSymChunk, r0 = (lane / kSymChunk) * kSymChunk;
          const int jl = cand[c0 + col] & (kSysTile - 1);
          const Atom4<S> aj = sj[jl];
          S sc = S(0), sx = S(0), sy = S(0), sz = S(0);
#pragma unroll 4
          for (int r = 0; r < kSymChunk; ++r) {
            const int row = r0 + r;
            const S c = S_[col * kSymPitch + row];
            sc += c;
            sx += c * Xi[row * 4]; sy += c * Xi[row * 4 + 1]; sz += c * Xi[row * 4 + 2];
          }
#pragma unroll
          for (int o = kSymChunk; o < 32; o <<= 1) { sc += shx(sc, o); sx += shx(sx, o); sy += shx(sy, o); sz += shx(sz, o); }
          if (lane < kSymChunk && (sc != S(0) || sx != S(0))) {
            const size_t j = (size_t)jt * kSysTile + jl;
            // grad_j += sum_i c (x_i - x_j)
            red_add(A.grad + 3 * j, sx - sc * aj.x);
            red_add(A.grad + 3 * j + 1, sy - sc * aj.y);
            red_add(A.grad + 3 * j + 2, sz - sc * aj.z);
          }
        }
        __syncwarp();
      }
    }
  }
  if (gx != S(0) || gy != S(0) || gz != S(0)) {
    red_add(A.grad + 3 * (size_t)i, gx); red_add(A.grad + 3 * (size_t)i + 1, gy); red_add(A.grad + 3 * (size_t)i + 2, gz);
  }
}


// =====================================================================================================
// CN sweeps on 8-row groups ("q8", round 2).  At the 25-Bohr CN cutoff a partner that can reach the ~13-Bohr box of a
// warp's 32 rows is within range of only ~40 % of them, so the 32-row form above spends most of its lanes on pairs that
// do not count.  Here a warp still owns the 32 rows of one sub-tile, but walks them as four groups of 8 rows: the lanes
// form four quarter-warps (lane = 8 q + r), quarter q takes its own partner from the group's candidate list and lane r
// the r-th row of the group.  Candidates are compacted against the group's ~8-Bohr box (lane efficiency ~53 %), the
// j-side sum of a partner is three xor-shuffles inside its quarter plus one atomic, the i-side sums stay in registers.
// The staged j-tile is shared by the 128 rows of the CTA exactly as before.
#ifndef D3_SYM_Q8
#define D3_SYM_Q8 1
#endif

// box of an 8-row group: min / max over the 8 lanes of a quarter (every lane of the quarter gets the result)
template <typename S>
__device__ __forceinline__ void sym_group_box(const Atom4<S>& a, S* gbox, int lane) {
  S lo[3] = {a.x, a.y, a.z}, hi[3] = {a.x, a.y, a.z};
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const S u = shx(lo[k], o), v = shx(hi[k], o);
      lo[k] = u < lo[k] ? u : lo[k];
      hi[k] = v > hi[k] ? v : hi[k];
    }
  }
  if ((lane & 7) == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { gbox[k] = lo[k]; gbox[3 + k] = hi[k]; }
  }
}

// candidates of one 8-row group in the staged j-tile (tile-local indices, in order, padded with 255 to a multiple of 4)
template <typename S>
__device__ __forceinline__ int sym_candidates_q8(const Atom4<S>* sj, const S* box, const S* gbox, int jt, int itile,
                                                 int wsub, S cut2, unsigned char* cand, int lane) {
  int n = 0;
#pragma unroll
  for (int sub = 0; sub < kSysTile / kSysSub; ++sub) {
    if (jt == itile && sub < wsub) continue;                                  // diagonal tile: only j > i counts
    if (box_dist2(gbox, box + ((size_t)jt * 4 + sub) * 6) > cut2) continue;   // warp-uniform
    const int jl = sub * kSysSub + lane;
    const bool ok = box_point_dist2(gbox, sj[jl]) <= cut2;
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (ok) cand[n + __popc(m & ((1u << lane) - 1))] = (unsigned char)jl;
    n += __popc(m);
  }
  if (lane < 4) cand[n + lane] = 255;
  __syncwarp();
  return n;
}

template <typename S>
__global__ void __launch_bounds__(kSysTile) sym_cn_q8_kernel(SymArgs<S> P) {
  const SysArgs<S>& A = P.sys;
  __shared__ Atom4<S> sj[kSysTile];
  __shared__ Atom4<S> si[kSysTile];                   // the CTA's own 128 rows
  __shared__ S sacc[kSysTile];                        // their CN sums
  __shared__ S sgbox[kSysTile / 8][6];                // boxes of the sixteen 8-row groups
  __shared__ unsigned char scand[kSysTile / 32][kSysTile + 8];
  int itile;
  if (!sym_row_tile(P, itile)) return;
  const int t = threadIdx.x, lane = t & 31, wsub = t >> 5, q = lane >> 3, r = lane & 7;
  unsigned char* cand = scand[wsub];
  {
    const Atom4<S> a = A.atoms[(size_t)itile * kSysTile + t];
    si[t] = a;
    sacc[t] = S(0);
    sym_group_box<S>(a, sgbox[t >> 3], lane);         // thread t holds row t: its quarter is group t / 8
  }
  const S tiny = tiny_of<S>();
  S ibox[6];
  tile_box(A.box, itile, ibox);
  const int ntiles = A.natoms / kSysTile;
  const int jt_first = itile + blockIdx.x;
  for (int k0 = 0; jt_first + k0 * A.jsplit < ntiles; k0 += 32)
  for (unsigned tmask = sym_tile_mask<S>(A.box, ibox, jt_first, A.jsplit, k0, ntiles, A.cn_cutoff2, lane); tmask;
       tmask &= tmask - 1) {
    const int jt = jt_first + (k0 + __ffs(tmask) - 1) * A.jsplit;      // CTA-uniform
    __syncthreads();
    sj[t] = A.atoms[(size_t)jt * kSysTile + t];
    __syncthreads();
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
      const int lrow = wsub * kSysSub + 8 * g + r;             // row of this lane in the group (tile-local)
      const int ncand = sym_candidates_q8<S>(sj, A.box, sgbox[wsub * 4 + g], jt, itile, wsub, A.cn_cutoff2, cand, lane);
      if (ncand == 0) continue;                                // warp-uniform
      const int i = itile * kSysTile + lrow;
      const Atom4<S> ai = si[lrow];
      S cn = S(0);
      for (int c0 = 0; c0 < ncand; c0 += 4) {
        const int jl = cand[c0 + q];                           // 255 = padding of the last step (quarter-uniform)
        S v = S(0);
        if (jl < kSysTile) {
          const int j = jt * kSysTile + jl;
          const Atom4<S> aj = sj[jl];
          S dx = ai.x - aj.x, dy = ai.y - aj.y, dz = ai.z - aj.z;
          S r2 = dx * dx + dy * dy + dz * dz;
          if (r2 <= A.cn_cutoff2 && j > i) {
            S rinv = frsq(r2 + tiny);
            S x = fmx(A.kcn - (ai.rk + aj.rk) * rinv, S(-700));
            v = frcp(S(1) + fex(x));
          }
        }
        cn += v;
        S s = v;
        s += shx(s, 1); s += shx(s, 2); s += shx(s, 4);         // sum over the 8 rows of the group, per quarter
        if (r == 0 && s != S(0)) red_add(A.cn + (size_t)jt * kSysTile + jl, s);
      }
      cn += shx(cn, 8); cn += shx(cn, 16);                     // the four quarters hold partial sums of the same rows
      if (q == 0) sacc[lrow] += cn;                            // (only this warp touches these rows)
      __syncwarp();                                            // `cand` is rebuilt for the next group
    }
  }
  __syncthreads();
  if (sacc[t] != S(0)) red_add(A.cn + (size_t)itile * kSysTile + t, sacc[t]);
}

template <typename S>
__global__ void __launch_bounds__(kSysTile) sym_cnbwd_q8_kernel(SymArgs<S> P) {
  const SysArgs<S>& A = P.sys;
  __shared__ Atom4<S> sj[kSysTile];
  __shared__ S sd[kSysTile];
  __shared__ Atom4<S> si[kSysTile];
  __shared__ S sdi[kSysTile];
  __shared__ S sacc[3][kSysTile];
  __shared__ S sgbox[kSysTile / 8][6];
  __shared__ unsigned char scand[kSysTile / 32][kSysTile + 8];
  int itile;
  if (!sym_row_tile(P, itile)) return;
  const int t = threadIdx.x, lane = t & 31, wsub = t >> 5, q = lane >> 3, r = lane & 7;
  unsigned char* cand = scand[wsub];
  {
    const Atom4<S> a = A.atoms[(size_t)itile * kSysTile + t];
    si[t] = a;
    sdi[t] = A.decn[(size_t)itile * kSysTile + t];
    sacc[0][t] = sacc[1][t] = sacc[2][t] = S(0);
    sym_group_box<S>(a, sgbox[t >> 3], lane);
  }
  const S tiny = tiny_of<S>(), eps = eps_of<S>();
  S ibox[6];
  tile_box(A.box, itile, ibox);
  const int ntiles = A.natoms / kSysTile;
  const int jt_first = itile + blockIdx.x;
  for (int k0 = 0; jt_first + k0 * A.jsplit < ntiles; k0 += 32)
  for (unsigned tmask = sym_tile_mask<S>(A.box, ibox, jt_first, A.jsplit, k0, ntiles, A.cn_cutoff2, lane); tmask;
       tmask &= tmask - 1) {
    const int jt = jt_first + (k0 + __ffs(tmask) - 1) * A.jsplit;      // CTA-uniform
    __syncthreads();
    sj[t] = A.atoms[(size_t)jt * kSysTile + t];
    sd[t] = A.decn[(size_t)jt * kSysTile + t];
    __syncthreads();
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
      const int lrow = wsub * kSysSub + 8 * g + r;
      const int ncand = sym_candidates_q8<S>(sj, A.box, sgbox[wsub * 4 + g], jt, itile, wsub, A.cn_cutoff2, cand, lane);
      if (ncand == 0) continue;
      const int i = itile * kSysTile + lrow;
      const Atom4<S> ai = si[lrow];
      const S dci = sdi[lrow];
      S gx = S(0), gy = S(0), gz = S(0);
      for (int c0 = 0; c0 < ncand; c0 += 4) {
        const int jl = cand[c0 + q];
        S cx = S(0), cy = S(0), cz = S(0);
        if (jl < kSysTile) {
          const int j = jt * kSysTile + jl;
          const Atom4<S> aj = sj[jl];
          S dx = ai.x - aj.x, dy = ai.y - aj.y, dz = ai.z - aj.z;
          S r2 = dx * dx + dy * dy + dz * dz;
          if (r2 <= A.cn_cutoff2 && r2 >= eps && j > i) {
            S rinv = frsq(r2 + tiny);
            S rk = ai.rk + aj.rk;
            S e = fex(fmx(A.kcn - rk * rinv, S(-700)));
            S f = frcp(S(1) + e);
            S df = (rk * (rinv * rinv * rinv)) * (f * (e * f));
            const S c = (dci + sd[jl]) * df;                   // grad_i -= c d, grad_j += c d
            cx = c * dx; cy = c * dy; cz = c * dz;
          }
        }
        gx -= cx; gy -= cy; gz -= cz;
        // sums over the 8 rows of the quarter, three values at once: a butterfly in which every level halves the
        // values a lane still carries (4 shuffles instead of 9); rows r = 0, 2, 4 end up with x, y, z
        {
          const bool up4 = (r & 4) != 0, up2 = (r & 2) != 0;
          S k0 = up4 ? cz : cx, k1 = up4 ? S(0) : cy;           // keep {x, y} (r < 4) or {z, 0}
          const S s0 = up4 ? cx : cz, s1 = up4 ? cy : S(0);     // send the other pair
          k0 += shx(s0, 4); k1 += shx(s1, 4);
          S k = up2 ? k1 : k0;                                    // keep x|z (r&2 == 0) or y|0
          k += shx(up2 ? k0 : k1, 2);
          k += shx(k, 1);
          if ((r & 1) == 0 && r < 6 && k != S(0) && jl < kSysTile)
            red_add(A.grad + 3 * ((size_t)jt * kSysTile + jl) + (r >> 1), k);
        }
      }
      gx += shx(gx, 8); gy += shx(gy, 8); gz += shx(gz, 8);
      gx += shx(gx, 16); gy += shx(gy, 16); gz += shx(gz, 16);
      if (q == 0) { sacc[0][lrow] += gx; sacc[1][lrow] += gy; sacc[2][lrow] += gz; }
      __syncwarp();
    }
  }
  __syncthreads();
  {
    const size_t i = (size_t)itile * kSysTile + t;
    const S x = sacc[0][t], y = sacc[1][t], z = sacc[2][t];
    if (x != S(0) || y != S(0) || z != S(0)) { red_add(A.grad + 3 * i, x); red_add(A.grad + 3 * i + 1, y); red_add(A.grad + 3 * i + 2, z); }
  }
}


// Sum v[0..7] over the 8 lanes of each quarter-warp: a butterfly in which every level halves the values a lane still
// carries (4 + 2 + 1 shuffles instead of 8 x 3); lane r of the quarter returns the total of v[r].
template <typename S>
__device__ __forceinline__ S quarter_reduce8(const S (&v)[8], int r) {
  const bool up4 = (r & 4) != 0, up2 = (r & 2) != 0, up1 = (r & 1) != 0;
  S k4[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) k4[m] = (up4 ? v[4 + m] : v[m]) + shx(up4 ? v[m] : v[4 + m], 4);
  S k2[2];
#pragma unroll
  for (int m = 0; m < 2; ++m) k2[m] = (up2 ? k4[2 + m] : k4[m]) + shx(up2 ? k4[m] : k4[2 + m], 2);
  return (up1 ? k2[1] : k2[0]) + shx(up1 ? k2[0] : k2[1], 1);
}

// ---------------------------------------------------------------- pair sweep on 8-row groups (T vectors from global)
template <typename S, bool WANT_G>
__global__ void __launch_bounds__(kSysTile) sym_pair_q8_kernel(SymArgs<S> P) {
  const SysArgs<S>& A = P.sys;
  const int E = A.nelem;
  __shared__ Atom4<S> sj[kSysTile];
  __shared__ Pair2<S> sb[kSysTile];
  __shared__ int sej[kSysTile];
  __shared__ Atom4<S> si[kSysTile];                   // the CTA's own 128 rows: position, {sqrt q, g}, element, w, dw
  __shared__ Pair2<S> sbi[kSysTile];
  __shared__ int sei[kSysTile];
  __shared__ __align__(16) S swi[kSysTile * 8];
  __shared__ __align__(16) S sdwi[WANT_G ? kSysTile * 8 : 8];
  __shared__ S sacc[WANT_G ? 5 : 1][kSysTile];        // i-side sums: E, gx, gy, gz, dL/dCN
  __shared__ S sgbox[kSysTile / 8][6];
  __shared__ unsigned char scand[kSysTile / 32][kSysTile + 8];
  int itile;
  if (!sym_row_tile(P, itile)) return;
  const int t = threadIdx.x, lane = t & 31, wsub = t >> 5, q = lane >> 3, r = lane & 7;
  unsigned char* cand = scand[wsub];
  {
    const size_t i = (size_t)itile * kSysTile + t;
    const Atom4<S> a = A.atoms[i];
    si[t] = a;
    sbi[t] = A.aux[i];
    sei[t] = A.eidx[i];
#pragma unroll
    for (int k = 0; k < 8; ++k) { swi[8 * t + k] = A.w[8 * i + k]; if (WANT_G) sdwi[8 * t + k] = A.dw[8 * i + k]; }
#pragma unroll
    for (int k = 0; k < (WANT_G ? 5 : 1); ++k) sacc[k][t] = S(0);
    sym_group_box<S>(a, sgbox[t >> 3], lane);
  }
  const S tiny = tiny_of<S>();
  const S* tdvec = A.tvec + (size_t)A.natoms * E * 8;
  S ibox[6];
  tile_box(A.box, itile, ibox);
  const int ntiles = A.natoms / kSysTile;
  const int jt_first = itile + blockIdx.x;
  for (int k0 = 0; jt_first + k0 * A.jsplit < ntiles; k0 += 32)
  for (unsigned tmask = sym_tile_mask<S>(A.box, ibox, jt_first, A.jsplit, k0, ntiles, A.cutoff2, lane); tmask;
       tmask &= tmask - 1) {
    const int jt = jt_first + (k0 + __ffs(tmask) - 1) * A.jsplit;      // CTA-uniform
    __syncthreads();
    {
      const size_t j = (size_t)jt * kSysTile + t;
      sj[t] = A.atoms[j];
      sb[t] = A.aux[j];
      sej[t] = A.eidx[j];
    }
    __syncthreads();
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
      const int lrow = wsub * kSysSub + 8 * g + r;
      const int ncand = sym_candidates_q8<S>(sj, A.box, sgbox[wsub * 4 + g], jt, itile, wsub, A.cutoff2, cand, lane);
      if (ncand == 0) continue;                                // warp-uniform
      const int i = itile * kSysTile + lrow;
      const Atom4<S> ai = si[lrow];
      const Pair2<S> bi = sbi[lrow];
      const int ei = sei[lrow];
      const int eic = ei < 0 ? 0 : ei;
      S wi[8], dwi[8];
      {
        const Pair2<S>* wp = reinterpret_cast<const Pair2<S>*>(swi + 8 * lrow);
        const Pair2<S>* dp = reinterpret_cast<const Pair2<S>*>(sdwi + (WANT_G ? 8 * lrow : 0));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const Pair2<S> p = wp[k]; wi[2 * k] = p.a; wi[2 * k + 1] = p.b;
          if (WANT_G) { const Pair2<S> d = dp[k]; dwi[2 * k] = d.a; dwi[2 * k + 1] = d.b; }
        }
      }
      const S gi = bi.b;
      const S s3qi = sqr(S(3)) * bi.a;
      S e_i = S(0), gx = S(0), gy = S(0), gz = S(0), decn = S(0);
      for (int c0 = 0; c0 < ncand; c0 += 4) {
        const int jlr = cand[c0 + q];                           // 255 = padding of the last step (quarter-uniform)
        const int jl = jlr & (kSysTile - 1);
        const int j = jt * kSysTile + jl;
        const int ej = jlr < kSysTile ? sej[jl] : -1;
        const Atom4<S> aj = sj[jl];
        S dx = ai.x - aj.x, dy = ai.y - aj.y, dz = ai.z - aj.z;
        S r2 = dx * dx + dy * dy + dz * dz;
        S v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = S(0);
        if (ej >= 0 && ei >= 0 && r2 <= A.cutoff2 && j > i) {
          r2 = r2 + tiny;
          const Pair2<S> bj = sb[jl];
          const Pair2<S>* tj = reinterpret_cast<const Pair2<S>*>(A.tvec + ((size_t)j * E + eic) * 8);
          S tv[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) { Pair2<S> p = tj[k]; tv[2 * k] = p.a; tv[2 * k + 1] = p.b; }
          S c6 = wi[0] * tv[0];
#pragma unroll
          for (int a = 1; a < kNRef; ++a) c6 += wi[a] * tv[a];
          S tq = s3qi * bj.a;
          S qq = tq * tq;
          S r0 = A.a1 * tq + A.a2;
          S r02 = r0 * r0;
          S r06 = r02 * r02 * r02;
          S r08 = r06 * r02;
          S r4 = r2 * r2;
          S r6 = r4 * r2;
          S r8 = r4 * r4;
          S d6 = r6 + r06, d8 = r8 + r08;
          S inv = frcp(d6 * d8);
          S t6 = inv * d8, t8 = inv * d6;
          S s8q = A.s8 * qq;
          S Pp = A.s6 * t6 + s8q * t8;
          const S ep = c6 * Pp;
          e_i += ep;
          v[0] = ep;
          if (WANT_G) {
            S dc6i = dwi[0] * tv[0];
#pragma unroll
            for (int a = 1; a < kNRef; ++a) dc6i += dwi[a] * tv[a];
            const Pair2<S>* tdj = reinterpret_cast<const Pair2<S>*>(tdvec + ((size_t)j * E + eic) * 8);
            S td[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) { Pair2<S> p = tdj[k]; td[2 * k] = p.a; td[2 * k + 1] = p.b; }
            S dc6j = wi[0] * td[0];
#pragma unroll
            for (int a = 1; a < kNRef; ++a) dc6j += wi[a] * td[a];
            S gam = S(0.5) * (gi + bj.b);
            S gP = gam * Pp;
            decn -= gP * dc6i;
            S t62 = t6 * t6, t82 = t8 * t8;
            S dP = S(3) * (A.s6 * (r4 * t62)) + S(4) * (s8q * (r6 * t82));
            const S fj = (S(2) * gam) * (c6 * dP);
            const S fx = fj * dx, fy = fj * dy, fz = fj * dz;
            gx += fx; gy += fy; gz += fz;
            v[1] = -fx; v[2] = -fy; v[3] = -fz;                 // grad_j = -sum_i f_ij (x_i - x_j)
            v[4] = -(gP * dc6j);
          }
        }
        if (WANT_G) {
          const S tot = quarter_reduce8<S>(v, r);               // lane r: total of v[r] over the 8 rows
          if (r < 5 && tot != S(0)) {
            S* dst = r == 0 ? A.energy + j : (r == 4 ? A.decn + j : A.grad + 3 * (size_t)j + (r - 1));
            red_add(dst, r == 0 ? S(-0.5) * tot : tot);
          }
        } else {
          S s = v[0];
          s += shx(s, 1); s += shx(s, 2); s += shx(s, 4);
          if (r == 0 && s != S(0)) red_add(A.energy + j, S(-0.5) * s);
        }
      }
      // the four quarters hold partial i-side sums of the same 8 rows
      e_i += shx(e_i, 8); e_i += shx(e_i, 16);
      if (WANT_G) {
        gx += shx(gx, 8); gy += shx(gy, 8); gz += shx(gz, 8); decn += shx(decn, 8);
        gx += shx(gx, 16); gy += shx(gy, 16); gz += shx(gz, 16); decn += shx(decn, 16);
      }
      if (q == 0) {
        sacc[0][lrow] += e_i;
        if (WANT_G) { sacc[1][lrow] += gx; sacc[2][lrow] += gy; sacc[3][lrow] += gz; sacc[4][lrow] += decn; }
      }
      __syncwarp();                                            // `cand` is rebuilt for the next group
    }
  }
  __syncthreads();
  {
    const size_t i = (size_t)itile * kSysTile + t;
    const S e_i = sacc[0][t];
    if (e_i != S(0)) {
      red_add(A.energy + i, S(-0.5) * e_i);
      if (WANT_G) {
        red_add(A.grad + 3 * i, sacc[1][t]); red_add(A.grad + 3 * i + 1, sacc[2][t]);
        red_add(A.grad + 3 * i + 2, sacc[3][t]); red_add(A.decn + i, sacc[4][t]);
      }
    }
  }
}

}  // namespace d3
