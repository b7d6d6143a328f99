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
constexpr int kSymChunk = D3_SYM_CHUNK;  // partners per transposed chunk
constexpr int kSymParts = 32 / kSymChunk; // lanes sharing one partner column
constexpr int kSymPitch = 33;            // 32 rows + 1: conflict-free column reads

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
__global__ void __launch_bounds__(kSysTile) sym_pair_t_kernel(SymArgs<S> P) {
  const SysArgs<S>& A = P.sys;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int E = A.nelem;
  Atom4<S>* sj = reinterpret_cast<Atom4<S>*>(smem_raw);
  Pair2<S>* sb = reinterpret_cast<Pair2<S>*>(sj + kSysTile);
  S* sT = reinterpret_cast<S*>(sb + kSysTile);                 // [128][E][8]  T vectors of the j tile
  S* sTd = sT + (size_t)kSysTile * E * 8;                      // [128][E][8]  Td vectors (dw-contracted)
  S* sSall = sTd + (size_t)kSysTile * E * 8;                   // [4 warps][3][chunk][33]
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
  unsigned char* cand = reinterpret_cast<unsigned char*>(sej + kSysTile) + wsub * (kSysTile + kSymChunk);
  Xi[lane * 4] = ai.x; Xi[lane * 4 + 1] = ai.y; Xi[lane * 4 + 2] = ai.z;
  __syncwarp();
  const int ntiles = A.natoms / kSysTile;
  const S gi = bi.b;
  const S s3qi = sqr(S(3)) * bi.a;
  S e_i = S(0), gx = S(0), gy = S(0), gz = S(0), decn = S(0);
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
      const Pair2<S>* src = reinterpret_cast<const Pair2<S>*>(A.tvec + j * E * 8);
      Pair2<S>* dst = reinterpret_cast<Pair2<S>*>(sT + (size_t)t * E * 8);
      for (int k = 0; k < 4 * E; ++k) dst[k] = src[k];
      if (WANT_G) {
        const Pair2<S>* srcd = reinterpret_cast<const Pair2<S>*>(tdvec + j * E * 8);
        Pair2<S>* dstd = reinterpret_cast<Pair2<S>*>(sTd + (size_t)t * E * 8);
        for (int k = 0; k < 4 * E; ++k) dstd[k] = srcd[k];
      }
    }
    __syncthreads();
    {
      const int ncand = sym_candidates<S>(sj, A.box, wbox, jt, itile, wsub, A.cutoff2, cand, lane);
      for (int c0 = 0; c0 < ncand; c0 += kSymChunk) {
#pragma unroll 2
        for (int jj = 0; jj < kSymChunk; ++jj) {
          const int jlr = cand[c0 + jj];                  // 255 = chunk padding (warp-uniform)
          const int jl = jlr & (kSysTile - 1);
          const int j = jt * kSysTile + jl;
          const int ej = jlr < kSysTile ? sej[jl] : -1;
          const Atom4<S> aj = sj[jl];
          S dx = ai.x - aj.x, dy = ai.y - aj.y, dz = ai.z - aj.z;
          S r2 = dx * dx + dy * dy + dz * dz;
          S ep = S(0), fj = S(0), dj = S(0);
          if (ej >= 0 && ei >= 0 && r2 <= A.cutoff2 && j > i) {
            r2 = r2 + tiny;
            const Pair2<S> bj = sb[jl];
            const Pair2<S>* tj = reinterpret_cast<const Pair2<S>*>(sT + ((size_t)jl * E + eic) * 8);
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
              const Pair2<S>* tdj = reinterpret_cast<const Pair2<S>*>(sTd + ((size_t)jl * E + eic) * 8);
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
          const Atom4<S> aj = sj[jl];
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
  return kSysTile * (sizeof(Atom4<S>) + sizeof(Pair2<S>) + sizeof(int)) +
         ((size_t)kSysTile * nelem * 16 + 4 * 3 * kSymChunk * kSymPitch + 4 * 32 * 4) * sizeof(S) +
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

}  // namespace d3
