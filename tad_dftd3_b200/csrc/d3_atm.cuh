// Axilrod-Teller-Muto term for one large structure (BASELINE config 5a: 10 k
// atoms with a triple cutoff), reference damping/atm.py:86-155.
//
// The reference sums 1/6 sum_j sum_k [r_ij<=c][r_jk<=c] e_ijk over dense N^3
// tensors.  e_ijk is symmetric, so with m_ab = [r_ab <= c] a triple contributes
// only if at least two of its three edges are short, and
//   L = sum_i g_i E_i = sum_{triples} w e,
//   w = 1/6 [g_i m_jk (m_ij+m_ik) + g_j m_ik (m_ij+m_jk) + g_k m_ij (m_ik+m_jk)].
//
// Enumeration: every SHORT edge (i,j), i < j, owned by one lane, walks the third
// atoms k with m_ik or m_jk.  A triple is thus met once per short edge (2 or 3
// times); each meeting accumulates only what belongs to the owner edge and its
// two atoms, in registers:
//   * F_ij += w de/d(r_ij^2),  G_ij += w e      (gradient / C6 chain of the edge)
//   * E_i += e/6 m_jk,  E_j += e/6 m_ik          (per-atom share / #meetings)
//   * if edge ik (jk) is long, its i-end (j-end) gradient and C6-chain terms --
//     a long edge is never an owner, its two ends are served by the two short
//     edges of the triple.
// A warp = (atom i, 32-atom sub-tile of partners j); k is warp-uniform and the
// k sub-tiles are culled by bounding boxes against i and against the j
// sub-tile.  Results are added with fp64 atomics once per (i, j) edge, not per
// triple.
#pragma once
#include "d3_system.cuh"

namespace d3 {

// e and its derivatives with respect to a = r_ij^2, b = r_ik^2, c = r_jk^2.
template <typename S, typename B>
__device__ __forceinline__ void atm_term3(S a, S b, S c, S c9, S r0, B pexp, bool need_b, bool need_c,
                                          S& e, S& de_a, S& de_b, S& de_c) {
  S r2 = a * b * c;
  S r1inv = frsq(r2);
  S r3inv = r1inv * r1inv * r1inv;
  S r5inv = r3inv * (r1inv * r1inv);
  S p = a + c - b, q = a - c + b, t = (c + b) - a;
  S s = p * q * t;
  S sr5 = s * r5inv;
  S ang = B(0.375) * sr5 + r3inv;
  S u = fex(fmx(pexp * lg(r0 * r1inv), B(-700)));   // (r0 / r1)^pexp
  S fd = frcp(B(1) + B(6) * u);
  e = c9 * (ang * fd);
  S dfd0 = (B(3) * pexp) * (u * (fd * fd));       // dfd/dx = dfd0 / x
  S qt = q * t, pt = p * t, pq = p * q;
  S k0 = c9 * fd, k1 = c9 * (ang * dfd0);
  // de/dx = k0 [0.375 (ds_x r5inv - 2.5 sr5 / x) - 1.5 r3inv / x] + k1 / x
  S common = k1 - k0 * (B(0.9375) * sr5 + B(1.5) * r3inv);   // coefficient of 1/x
  S k0r5 = B(0.375) * (k0 * r5inv);
  de_a = k0r5 * (qt + pt - pq) + common * frcp(a);
  de_b = need_b ? k0r5 * (pt + pq - qt) + common * frcp(b) : S(0);
  de_c = need_c ? k0r5 * (qt + pq - pt) + common * frcp(c) : S(0);
}

template <typename S>
struct AtmArgs {
  SysArgs<S> sys;          // atoms, aux (g), z, box, w, dw, refc6, cutoff2, energy/grad/decn outputs
  const S* rvdw;           // [104][104]
  S s9, rs9, pexp;
};

template <typename S>
__device__ __forceinline__ void contract_k(const S* __restrict__ K, const S* w, const S* dw, S* U, S* Ud) {
#pragma unroll
  for (int b = 0; b < kNRef; ++b) { U[b] = S(0); Ud[b] = S(0); }
#pragma unroll
  for (int a = 0; a < kNRef; ++a) {
#pragma unroll
    for (int b = 0; b < kNRef; ++b) {
      const S k = __ldg(K + a * kNRef + b);
      U[b] += w[a] * k;
      Ud[b] += dw[a] * k;
    }
  }
}

__device__ __forceinline__ void red_add(double* p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(float* p, float v) { atomicAdd(p, v); }

constexpr int kAtmWarps = 4;

template <typename S, bool WANT_G>
__global__ void __launch_bounds__(32 * kAtmWarps) system_atm_kernel(AtmArgs<S> P) {
  const SysArgs<S>& A = P.sys;
  struct KTile {
    Atom4<S> at[kSysSub];
    S g[kSysSub];
    __align__(16) S w[kSysSub * 8];
    int z[kSysSub];
  };
  __shared__ KTile tiles[kAtmWarps];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  KTile& kt = tiles[wid];
  const int i = A.row_begin + blockIdx.x;
  const int Zi = A.z[i];
  if (Zi == 0) return;                                  // padding row (whole CTA)
  const Atom4<S> ai = A.atoms[i];
  const S gi = A.aux[i].b;
  S wi[kNRef], dwi[kNRef];
#pragma unroll
  for (int a = 0; a < kNRef; ++a) { wi[a] = A.w[8 * (size_t)i + a]; dwi[a] = A.dw[8 * (size_t)i + a]; }
  const S ibox[6] = {ai.x, ai.y, ai.z, ai.x, ai.y, ai.z};
  const int nsub = A.natoms / kSysSub;
  const S sixth = S(1) / S(6);
  const S rs93 = P.rs9 * P.rs9 * P.rs9;
  const S tiny = tiny_of<S>();

  for (int js = (i / kSysSub) + wid; js < nsub; js += kAtmWarps) {     // partners j > i only
    const S* jbox = A.box + (size_t)js * 6;
    if (box_dist2(ibox, jbox) > A.cutoff2) continue;                  // warp-uniform
    const int j = js * kSysSub + lane;
    const int Zj = A.z[j];
    const Atom4<S> aj = A.atoms[j];
    S dxij = ai.x - aj.x, dyij = ai.y - aj.y, dzij = ai.z - aj.z;
    S a2 = dxij * dxij + dyij * dyij + dzij * dzij;
    const bool edge = Zj != 0 && j > i && a2 <= A.cutoff2;            // this lane owns a short edge
    if (!__any_sync(0xffffffffu, edge)) continue;
    a2 = a2 + tiny;
    const S gj = A.aux[j].b;
    S wj[kNRef], dwj[kNRef];
#pragma unroll
    for (int a = 0; a < kNRef; ++a) { wj[a] = A.w[8 * (size_t)j + a]; dwj[a] = A.dw[8 * (size_t)j + a]; }
    const int Zjc = Zj ? Zj : Zi;
    // edge constants: C6_ij and its CN derivatives
    S c6ij, dc6_i, dc6_j;
    {
      S U[kNRef], Ud[kNRef];
      contract_k<S>(A.refc6 + ((size_t)Zi * kMaxZ + Zjc) * 49, wi, dwi, U, Ud);
      c6ij = S(0); dc6_i = S(0); dc6_j = S(0);
#pragma unroll
      for (int b = 0; b < kNRef; ++b) { c6ij += U[b] * wj[b]; dc6_i += Ud[b] * wj[b]; dc6_j += U[b] * dwj[b]; }
    }
    const S hij = sqr(ab(c6ij) + tiny);
    const S Rij = P.rvdw[Zi * kMaxZ + Zjc];
    S Fij = S(0), Gij = S(0), Ei = S(0), Ej = S(0);
    S gix = S(0), giy = S(0), giz = S(0), gjx = S(0), gjy = S(0), gjz = S(0), dci = S(0), dcj = S(0);
    S Ui[kNRef], Udi[kNRef], Uj[kNRef], Udj[kNRef];
    S Rik = S(0), Rjk = S(0);
    int Zprev = -1;

    for (int ks = 0; ks < nsub; ++ks) {
      const S* kbox = A.box + (size_t)ks * 6;
      const bool near_i = box_dist2(ibox, kbox) <= A.cutoff2;
      const bool near_j = box_dist2(jbox, kbox) <= A.cutoff2;
      if (!(near_i || near_j)) continue;                              // warp-uniform
      __syncwarp();
      {
        const size_t k = (size_t)ks * kSysSub + lane;
        kt.at[lane] = A.atoms[k];
        kt.g[lane] = A.aux[k].b;
        kt.z[lane] = A.z[k];
        const Pair2<S>* src = reinterpret_cast<const Pair2<S>*>(A.w + 8 * k);
        Pair2<S>* dst = reinterpret_cast<Pair2<S>*>(kt.w + 8 * lane);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = src[q];
      }
      __syncwarp();
      for (int kk = 0; kk < kSysSub; ++kk) {
        const int Zk = kt.z[kk];
        const int k = ks * kSysSub + kk;
        if (Zk == 0 || k == i) continue;                              // warp-uniform
        const Atom4<S> ak = kt.at[kk];
        S dxik = ai.x - ak.x, dyik = ai.y - ak.y, dzik = ai.z - ak.z;
        S b2 = dxik * dxik + dyik * dyik + dzik * dzik;
        const bool mik = b2 <= A.cutoff2;                             // warp-uniform
        S dxjk = aj.x - ak.x, dyjk = aj.y - ak.y, dzjk = aj.z - ak.z;
        S c2 = dxjk * dxjk + dyjk * dyjk + dzjk * dzjk;
        const bool mjk = c2 <= A.cutoff2;
        const bool act = edge && k != j && (mik || mjk);
        if (!__any_sync(0xffffffffu, act)) continue;
        if (Zk != Zprev) {
          Zprev = Zk;
          contract_k<S>(A.refc6 + ((size_t)Zi * kMaxZ + Zk) * 49, wi, dwi, Ui, Udi);
          contract_k<S>(A.refc6 + ((size_t)Zjc * kMaxZ + Zk) * 49, wj, dwj, Uj, Udj);
          Rik = P.rvdw[Zi * kMaxZ + Zk];
          Rjk = P.rvdw[Zjc * kMaxZ + Zk];
        }
        if (act) {
          b2 = b2 + tiny; c2 = c2 + tiny;
          const Pair2<S>* wk2 = reinterpret_cast<const Pair2<S>*>(kt.w + 8 * kk);
          S wk[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) { Pair2<S> pp = wk2[q]; wk[2 * q] = pp.a; wk[2 * q + 1] = pp.b; }
          S c6ik = S(0), c6jk = S(0);
#pragma unroll
          for (int b = 0; b < kNRef; ++b) { c6ik += Ui[b] * wk[b]; c6jk += Uj[b] * wk[b]; }
          S c9 = P.s9 * (hij * sqr(ab(c6ik * c6jk) + tiny));
          S r0 = rs93 * (Rij * Rik * Rjk);
          S e, dea, deb, dec;
          atm_term3<S, S>(a2, b2, c2, c9, r0, P.pexp, !mik, !mjk, e, dea, deb, dec);
          const S fik = mik ? S(1) : S(0), fjk = mjk ? S(1) : S(0);
          Ei += e * fjk;                       // share_i / #meetings that own i = e/6 m_jk
          Ej += e * fik;
          if (WANT_G) {
            // 6 w = g_i m_jk (1 + m_ik) + g_j m_ik (1 + m_jk) + g_k (m_ik + m_jk)
            S w6 = gi * (fjk * (S(1) + fik)) + gj * (fik * (S(1) + fjk)) + kt.g[kk] * (fik + fjk);
            S we = w6 * e;
            Fij += w6 * dea;
            Gij += we;
            if (!mik) {          // long edge ik: its i-end belongs to this meeting
              S f = w6 * deb;
              gix += f * dxik; giy += f * dyik; giz += f * dzik;
              S d = S(0);
#pragma unroll
              for (int b = 0; b < kNRef; ++b) d += Udi[b] * wk[b];
              dci += we * (d * rcp0(c6ik));
            }
            if (!mjk) {          // long edge jk: its j-end
              S f = w6 * dec;
              gjx += f * dxjk; gjy += f * dyjk; gjz += f * dzjk;
              S d = S(0);
#pragma unroll
              for (int b = 0; b < kNRef; ++b) d += Udj[b] * wk[b];
              dcj += we * (d * rcp0(c6jk));
            }
          }
        }
      }
    }
    // ---- finish the edge: scale, add the owner-edge terms, write out
    S ei = sixth * Ei, ej = sixth * Ej;
    if (WANT_G) {
      const S two6 = S(2) * sixth, half6 = S(0.5) * sixth;
      S f = two6 * Fij;
      gix = two6 * gix + f * dxij; giy = two6 * giy + f * dyij; giz = two6 * giz + f * dzij;
      gjx = two6 * gjx - f * dxij; gjy = two6 * gjy - f * dyij; gjz = two6 * gjz - f * dzij;
      S gc = half6 * (Gij * rcp0(c6ij));
      dci = half6 * dci + gc * dc6_i;
      dcj = half6 * dcj + gc * dc6_j;
    }
    if (!edge) { ei = ej = S(0); gix = giy = giz = gjx = gjy = gjz = dci = dcj = S(0); }
    // j side: one atomic per quantity and lane
    if (edge) {
      red_add(A.energy + j, ej);
      if (WANT_G) {
        red_add(A.grad + 3 * (size_t)j, gjx); red_add(A.grad + 3 * (size_t)j + 1, gjy);
        red_add(A.grad + 3 * (size_t)j + 2, gjz); red_add(A.decn + j, dcj);
      }
    }
    // i side: warp sum, then one atomic per quantity
    ei = warp_sum<S>(ei);
    if (WANT_G) { gix = warp_sum<S>(gix); giy = warp_sum<S>(giy); giz = warp_sum<S>(giz); dci = warp_sum<S>(dci); }
    if (lane == 0) {
      red_add(A.energy + i, ei);
      if (WANT_G) {
        red_add(A.grad + 3 * (size_t)i, gix); red_add(A.grad + 3 * (size_t)i + 1, giy);
        red_add(A.grad + 3 * (size_t)i + 2, giz); red_add(A.decn + i, dci);
      }
    }
  }
}


// ===========================================================================
// Second-generation ATM kernel for structures with at most kAtmEmax elements
// (profile of the first one on 3000 atoms: 254 registers -> 8 warps/SM, half
// of the lanes without an owned edge, library log/exp in the damping power).
//
//  * per-atom pre-contracted vectors T_k[e][a] = sum_b K[Z_e, Z_k, a, b] w_kb
//    (one tiny kernel per call) replace the in-loop 7x7 contractions: every C6
//    and dC6/dCN the kernel needs is a 7-wide dot of the lane's own w or dw
//    with T_k[e_lane];
//  * the partners j > i of atom i are compacted into a shared list, so every
//    lane of a batch owns a short edge;
//  * (r0/r)^(16/3) for the default alp = 14 uses a MUFU-seeded inverse cube
//    root instead of exp(log).
// ===========================================================================
constexpr int kAtmEmax = 8;
constexpr int kAtmList = 1024;           // partners gathered per round (32 sub-tiles)

template <typename S>
struct AtmArgs2 {
  SysArgs<S> sys;
  const S* tvec;            // = sys.tvec  [natoms][nelem][8]  (7 used)
  const signed char* eidx;  // = sys.eidx  compact element index (-1 = padding)
  const S* rvdw_e;          // [nelem][nelem] compact element-pair radii
  int nelem;
  S s9, rs9, pexp;
  int cube_root_pow;        // 1: pexp == 16/3
};

// T vectors of every atom (grid over atoms, one thread per (atom, element))
// tvec  [natoms][nelem][8]: T_k[e][a]  = sum_b K[Z_e, Z_k, a, b] w_kb
// tdvec [natoms][nelem][8]: Td_k[e][a] = sum_b K[Z_e, Z_k, a, b] dw_kb   (right behind tvec)
template <typename S>
__global__ void system_tvec_kernel(int natoms, int nelem, const int* z, const int* zlist, const S* w,
                                   const S* dw, const S* refc6, S* tvec) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= natoms * nelem) return;
  const int k = idx / nelem, e = idx % nelem;
  const int Zk = z[k], Ze = zlist[e];
  S out[8], outd[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) { out[a] = S(0); outd[a] = S(0); }
  if (Zk > 0) {
    const S* K = refc6 + ((size_t)Ze * kMaxZ + Zk) * 49;
#pragma unroll
    for (int a = 0; a < kNRef; ++a) {
#pragma unroll
      for (int b = 0; b < kNRef; ++b) {
        const S kab = __ldg(K + a * kNRef + b);
        out[a] += kab * w[8 * (size_t)k + b];
        outd[a] += kab * dw[8 * (size_t)k + b];
      }
    }
  }
  S* tdvec = tvec + (size_t)natoms * nelem * 8;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    tvec[((size_t)k * nelem + e) * 8 + a] = out[a];
    tdvec[((size_t)k * nelem + e) * 8 + a] = outd[a];
  }
}

// x^(16/3) for x > 0: y = x^(-1/3) from a float seed and one cubic step, then (x y^2)^16
template <typename S>
__device__ __forceinline__ S pow16_3(S x) {
  S y = S(exp2f(-0.33333334f * log2f((float)x)));
  S y3 = y * y * y;
  S e = S(1) - x * y3;
  y = y + (y * e) * (S(1.0 / 3.0) + S(2.0 / 9.0) * e);
  S c = x * (y * y);          // x^(1/3)
  S c2 = c * c, c4 = c2 * c2, c8 = c4 * c4;
  return c8 * c8;
}

template <typename S, bool WANT_G>
__global__ void __launch_bounds__(32 * kAtmWarps, 4) system_atm2_kernel(AtmArgs2<S> P) {
  const SysArgs<S>& A = P.sys;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int E = P.nelem;
  // layout: jlist | count | rv[E*E] | per-warp tiles {at[32] g[32] e[32] T[32*E*8]}
  int* jlist = reinterpret_cast<int*>(smem_raw);
  int* jcount = jlist + kAtmList;
  S* rv = reinterpret_cast<S*>(jcount + 4);
  S* wiS = rv + kAtmEmax * kAtmEmax;                 // w_i[8], dw_i[8]
  unsigned char* tile_base = reinterpret_cast<unsigned char*>(wiS + 16);
  const size_t tile_bytes = kSysSub * (sizeof(Atom4<S>) + sizeof(S) + 8) + (size_t)kSysSub * E * 8 * sizeof(S);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned char* mytile = tile_base + wid * ((tile_bytes + 15) / 16 * 16);
  Atom4<S>* tat = reinterpret_cast<Atom4<S>*>(mytile);
  S* tg = reinterpret_cast<S*>(tat + kSysSub);
  S* tT = tg + kSysSub;                              // [32][E][8]
  int* te = reinterpret_cast<int*>(tT + (size_t)kSysSub * E * 8);

  const int i = A.row_begin + blockIdx.x;
  const int ei = P.eidx[i];
  if (ei < 0) return;                                // padding row (whole CTA)
  const Atom4<S> ai = A.atoms[i];
  const S gi = A.aux[i].b;
  for (int k = threadIdx.x; k < E * E; k += blockDim.x) rv[k] = P.rvdw_e[k];
  if (threadIdx.x < 8) { wiS[threadIdx.x] = A.w[8 * (size_t)i + threadIdx.x]; wiS[8 + threadIdx.x] = A.dw[8 * (size_t)i + threadIdx.x]; }
  const S ibox[6] = {ai.x, ai.y, ai.z, ai.x, ai.y, ai.z};
  const int nsub = A.natoms / kSysSub;
  const S sixth = S(1) / S(6);
  const S rs93 = P.rs9 * P.rs9 * P.rs9;
  const S tiny = tiny_of<S>();

  for (int js0 = i / kSysSub; js0 < nsub; js0 += kAtmList / kSysSub) {
    // ---- gather the partners j > i within the cutoff from up to 32 sub-tiles
    __syncthreads();
    if (threadIdx.x == 0) *jcount = 0;
    __syncthreads();
    for (int js = js0 + wid; js < min(js0 + kAtmList / kSysSub, nsub); js += kAtmWarps) {
      if (box_dist2(ibox, A.box + (size_t)js * 6) > A.cutoff2) continue;
      const int j = js * kSysSub + lane;
      const Atom4<S> aj = A.atoms[j];
      S dx = ai.x - aj.x, dy = ai.y - aj.y, dz = ai.z - aj.z;
      const bool edge = P.eidx[j] >= 0 && j > i && (dx * dx + dy * dy + dz * dz) <= A.cutoff2;
      const unsigned m = __ballot_sync(0xffffffffu, edge);
      int base = 0;
      if (lane == 0 && m) base = atomicAdd(jcount, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (edge) jlist[base + __popc(m & ((1u << lane) - 1))] = j;
    }
    __syncthreads();
    const int nj = *jcount;
    // ---- batches of 32 owner edges
    for (int b0 = wid * 32; b0 < nj; b0 += 32 * kAtmWarps) {
      const bool edge = b0 + lane < nj;
      const int j = jlist[edge ? b0 + lane : b0];
      const int ej = P.eidx[j];
      const Atom4<S> aj = A.atoms[j];
      const S gj = A.aux[j].b;
      S dxij = ai.x - aj.x, dyij = ai.y - aj.y, dzij = ai.z - aj.z;
      const S a2 = dxij * dxij + dyij * dyij + dzij * dzij + tiny;
      S wj[kNRef], dwj[kNRef];
#pragma unroll
      for (int a = 0; a < kNRef; ++a) { wj[a] = A.w[8 * (size_t)j + a]; dwj[a] = A.dw[8 * (size_t)j + a]; }
      // edge constants from T_j[e_i]
      S c6ij = S(0), dc6_i = S(0), dc6_j = S(0);
      {
        const S* Tj = P.tvec + ((size_t)j * E + ei) * 8;     // sum_b K[Zi,Zj,a,b] w_jb
        const S* Ti = P.tvec + ((size_t)i * E + ej) * 8;     // sum_b K[Zj,Zi,a,b] w_ib
#pragma unroll
        for (int a = 0; a < kNRef; ++a) {
          c6ij += wiS[a] * Tj[a];
          dc6_i += wiS[8 + a] * Tj[a];
          dc6_j += dwj[a] * Ti[a];
        }
      }
      const S hij = sqr(ab(c6ij) + tiny);
      const S Rij = rv[ei * E + ej];
      // bounding box of the batch
      S jbox[6] = {aj.x, aj.y, aj.z, aj.x, aj.y, aj.z};
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          S lo = shx(jbox[q], o), hi = shx(jbox[3 + q], o);
          jbox[q] = lo < jbox[q] ? lo : jbox[q];
          jbox[3 + q] = hi > jbox[3 + q] ? hi : jbox[3 + q];
        }
      }
      S Fij = S(0), Gij = S(0), Ei = S(0), Ej = S(0);
      S gix = S(0), giy = S(0), giz = S(0), gjx = S(0), gjy = S(0), gjz = S(0), dci = S(0), dcj = S(0);

      for (int ks = 0; ks < nsub; ++ks) {
        const S* kbox = A.box + (size_t)ks * 6;
        if (!(box_dist2(ibox, kbox) <= A.cutoff2 || box_dist2(jbox, kbox) <= A.cutoff2)) continue;
        __syncwarp();
        {
          const size_t k = (size_t)ks * kSysSub + lane;
          tat[lane] = A.atoms[k];
          tg[lane] = A.aux[k].b;
          te[lane] = P.eidx[k];
          const Pair2<S>* src = reinterpret_cast<const Pair2<S>*>(P.tvec + k * E * 8);
          Pair2<S>* dst = reinterpret_cast<Pair2<S>*>(tT + (size_t)lane * E * 8);
          for (int q = 0; q < 4 * E; ++q) dst[q] = src[q];
        }
        __syncwarp();
        for (int kk = 0; kk < kSysSub; ++kk) {
          const int ek = te[kk];
          const int k = ks * kSysSub + kk;
          if (ek < 0 || k == i) continue;                              // warp-uniform
          const Atom4<S> ak = tat[kk];
          S dxik = ai.x - ak.x, dyik = ai.y - ak.y, dzik = ai.z - ak.z;
          S b2 = dxik * dxik + dyik * dyik + dzik * dzik;
          const bool mik = b2 <= A.cutoff2;                             // warp-uniform
          S dxjk = aj.x - ak.x, dyjk = aj.y - ak.y, dzjk = aj.z - ak.z;
          S c2 = dxjk * dxjk + dyjk * dyjk + dzjk * dzjk;
          const bool mjk = c2 <= A.cutoff2;
          const bool act = edge && k != j && (mik || mjk);
          if (!__any_sync(0xffffffffu, act)) continue;
          if (act) {
            b2 = b2 + tiny; c2 = c2 + tiny;
            const S* Tki = tT + ((size_t)kk * E + ei) * 8;              // K[Zi,Zk] w_k
            const S* Tkj = tT + ((size_t)kk * E + ej) * 8;              // K[Zj,Zk] w_k
            S c6ik = S(0), c6jk = S(0);
#pragma unroll
            for (int a = 0; a < kNRef; ++a) { c6ik += wiS[a] * Tki[a]; c6jk += wj[a] * Tkj[a]; }
            S c9 = P.s9 * (hij * sqr(ab(c6ik * c6jk) + tiny));
            S r0 = rs93 * (Rij * rv[ei * E + ek] * rv[ej * E + ek]);
            // e and derivatives (atm_term3 with the power chosen at run time)
            S r2 = a2 * b2 * c2;
            S r1inv = frsq(r2);
            S r3inv = r1inv * r1inv * r1inv;
            S r5inv = r3inv * (r1inv * r1inv);
            S p = a2 + c2 - b2, q = a2 - c2 + b2, t = (c2 + b2) - a2;
            S s = p * q * t;
            S sr5 = s * r5inv;
            S ang = S(0.375) * sr5 + r3inv;
            S base = r0 * r1inv;
            S u = P.cube_root_pow ? pow16_3<S>(base) : fex(fmx(P.pexp * lg(base), S(-700)));
            S fd = frcp(S(1) + S(6) * u);
            S e = c9 * (ang * fd);
            const S fik = mik ? S(1) : S(0), fjk = mjk ? S(1) : S(0);
            Ei += e * fjk;
            Ej += e * fik;
            if (WANT_G) {
              S dfd0 = (S(3) * P.pexp) * (u * (fd * fd));
              S qt = q * t, pt = p * t, pq = p * q;
              S k0 = c9 * fd, k1 = c9 * (ang * dfd0);
              S common = k1 - k0 * (S(0.9375) * sr5 + S(1.5) * r3inv);
              S k0r5 = S(0.375) * (k0 * r5inv);
              S dea = k0r5 * (qt + pt - pq) + common * frcp(a2);
              S w6 = gi * (fjk * (S(1) + fik)) + gj * (fik * (S(1) + fjk)) + tg[kk] * (fik + fjk);
              S we = w6 * e;
              Fij += w6 * dea;
              Gij += we;
              if (!mik) {
                S f = w6 * (k0r5 * (pt + pq - qt) + common * frcp(b2));
                gix += f * dxik; giy += f * dyik; giz += f * dzik;
                S d = S(0);
#pragma unroll
                for (int a = 0; a < kNRef; ++a) d += wiS[8 + a] * Tki[a];
                dci += we * (d * rcp0(c6ik));
              }
              if (!mjk) {
                S f = w6 * (k0r5 * (qt + pq - pt) + common * frcp(c2));
                gjx += f * dxjk; gjy += f * dyjk; gjz += f * dzjk;
                S d = S(0);
#pragma unroll
                for (int a = 0; a < kNRef; ++a) d += dwj[a] * Tkj[a];
                dcj += we * (d * rcp0(c6jk));
              }
            }
          }
        }
      }
      // ---- finish the batch
      S ei_ = sixth * Ei, ej_ = sixth * Ej;
      if (WANT_G) {
        const S two6 = S(2) * sixth, half6 = S(0.5) * sixth;
        S f = two6 * Fij;
        gix = two6 * gix + f * dxij; giy = two6 * giy + f * dyij; giz = two6 * giz + f * dzij;
        gjx = two6 * gjx - f * dxij; gjy = two6 * gjy - f * dyij; gjz = two6 * gjz - f * dzij;
        S gc = half6 * (Gij * rcp0(c6ij));
        dci = half6 * dci + gc * dc6_i;
        dcj = half6 * dcj + gc * dc6_j;
      }
      if (!edge) { ei_ = ej_ = S(0); gix = giy = giz = gjx = gjy = gjz = dci = dcj = S(0); }
      if (edge) {
        red_add(A.energy + j, ej_);
        if (WANT_G) {
          red_add(A.grad + 3 * (size_t)j, gjx); red_add(A.grad + 3 * (size_t)j + 1, gjy);
          red_add(A.grad + 3 * (size_t)j + 2, gjz); red_add(A.decn + j, dcj);
        }
      }
      ei_ = warp_sum<S>(ei_);
      if (WANT_G) { gix = warp_sum<S>(gix); giy = warp_sum<S>(giy); giz = warp_sum<S>(giz); dci = warp_sum<S>(dci); }
      if (lane == 0) {
        red_add(A.energy + i, ei_);
        if (WANT_G) {
          red_add(A.grad + 3 * (size_t)i, gix); red_add(A.grad + 3 * (size_t)i + 1, giy);
          red_add(A.grad + 3 * (size_t)i + 2, giz); red_add(A.decn + i, dci);
        }
      }
    }
  }
}

template <typename S>
__global__ void system_rvdw_compact_kernel(int nelem, const int* zlist, const S* rvdw, S* rvdw_e) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nelem * nelem) return;
  rvdw_e[k] = rvdw[zlist[k / nelem] * kMaxZ + zlist[k % nelem]];
}

template <typename S>
inline size_t atm2_smem_bytes(int nelem) {
  size_t tile = kSysSub * (sizeof(Atom4<S>) + sizeof(S) + 8) + (size_t)kSysSub * nelem * 8 * sizeof(S);
  tile = (tile + 15) / 16 * 16;
  return (kAtmList + 4) * sizeof(int) + (kAtmEmax * kAtmEmax + 16) * sizeof(S) + kAtmWarps * tile + 64;
}

}  // namespace d3
