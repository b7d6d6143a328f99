// Third-generation ATM kernel: every triple is evaluated ONCE.
//
// Reference: damping/atm.py:86-155 (ordered cutoff rule: the mask tests r_ij and
// r_jk twice, never r_ik, atm.py:147-149).  With m_ab = [r_ab <= c] a triple
// contributes iff at least two of its edges are short, i.e. iff it has a CENTRE:
// a vertex whose two edges are short.  A path a-b-c (ac long) has exactly one
// centre; a triangle has three, and is assigned to its smallest index.  For the
// centre i and the pair (j, k), m_ij = m_ik = 1 and with f = m_jk
//   E_i += e/6 * 2f,   E_j += e/6 * (1 + f),   E_k += e/6 * (1 + f)
// which is the reference's ordered rule written per unordered triple.
// (The second generation, `system_atm2_kernel`, meets a triple once per short
// edge, 2.3 times on average for BASELINE config 5a.)
//
// Work decomposition
//  * persistent CTAs take centres i from a global counter (a launch owns the
//    centres row_begin..row_end-1, which is also how ranks shard the term);
//  * the CTA compacts the neighbours of i (r <= c), in index order, into its own
//    list in global scratch, with one record per neighbour n:
//    {sqrt|C6_in|, r_in^2, 1/r_in^2, (dC6_in/dCN_i)/C6_in, (dC6_in/dCN_n)/C6_in};
//  * warps take 32-entry tiles of the list (largest first): lane <-> k, and walk
//    all list positions j before k (j is warp-uniform).  Everything that belongs
//    to k or to edge ik accumulates in registers over the j loop; the j-side sums
//    of a step (7 values) go through a transposed shared-memory reduction and are
//    added to the global arrays with five fp64 atomics per 32 triples; the i-side
//    sums stay in registers until the centre is finished;
//  * C6_jk and its two CN derivatives are 7-wide dots of the lane's w_k / dw_k
//    with the pre-contracted vectors T_j[e_k], Td_j[e_k] (`system_tvec_kernel`).
#pragma once
#include "d3_atm.cuh"

namespace d3 {

constexpr int kAtm3Warps = 8;
constexpr int kAtm3Ctas = 2;         // resident CTAs per SM the kernel is built for
constexpr int kAtm3Rec = 6;          // reals per neighbour record (5 used)
constexpr int kAtm3RedStride = 40;   // row stride of the reduction tile (rows 16 banks apart)

template <typename S>
struct AtmArgs3 {
  AtmArgs2<S> a;
  int* list;       // [gridDim.x][natoms]          neighbour indices of the CTA's current centre
  S* rec;          // [gridDim.x][natoms][kAtm3Rec] neighbour records
  int* counter;    // zero before the launch
};

template <typename S>
inline size_t atm3_workspace_bytes(int natoms, int grid) {
  return 256 + (size_t)grid * natoms * (sizeof(int) + kAtm3Rec * sizeof(S));
}

// Column sums of vals[V] over the 32 lanes through this warp's tile [V][kAtm3RedStride].
// On return lanes 4r .. 4r+3 hold the total of row r (r < V <= 8).
template <int V, typename S>
__device__ __forceinline__ S warp_rows_reduce(const S (&vals)[V], S* tile, int lane) {
  static_assert(V <= 8, "at most 8 rows");
#pragma unroll
  for (int v = 0; v < V; ++v) tile[v * kAtm3RedStride + lane] = vals[v];
  __syncwarp();
  const int r = lane >> 2, part = lane & 3;
  S tot = S(0);
  if (r < V) {
    const Pair2<S>* rp = reinterpret_cast<const Pair2<S>*>(tile + r * kAtm3RedStride);
    S s0 = S(0), s1 = S(0);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const Pair2<S> p = rp[part + 4 * k]; s0 += p.a; s1 += p.b; }
    tot = s0 + s1;
  }
  tot += shx(tot, 1);
  tot += shx(tot, 2);
  __syncwarp();
  return tot;
}

__device__ __forceinline__ float bcast(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

template <typename S, bool WANT_G>
__global__ void __launch_bounds__(32 * kAtm3Warps, kAtm3Ctas) system_atm3_kernel(AtmArgs3<S> Q) {
  const AtmArgs2<S>& P = Q.a;
  const SysArgs<S>& A = P.sys;
  constexpr int nw = kAtm3Warps;
  constexpr int NV = WANT_G ? 7 : 1;
  __shared__ S rv[kAtmEmax * kAtmEmax];
  __shared__ __align__(16) S wiS[16];
  __shared__ __align__(16) S red[nw * 7 * kAtm3RedStride];
  __shared__ S isum[nw][5];
  __shared__ int cnt[nw];
  __shared__ int s_center, s_tile;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int E = P.nelem;
  const int natoms = A.natoms;
  const int nsub = natoms / kSysSub;
  const int nrows = A.row_end - A.row_begin;
  int* list = Q.list + (size_t)blockIdx.x * natoms;
  S* rec = Q.rec + (size_t)blockIdx.x * natoms * kAtm3Rec;
  S* redw = red + wid * 7 * kAtm3RedStride;
  const S* tdvec = P.tvec + (size_t)natoms * E * 8;
  const S sixth = S(1) / S(6), two6 = S(2) * sixth, half6 = S(0.5) * sixth;
  const S rs93 = P.rs9 * P.rs9 * P.rs9;
  const S tiny = tiny_of<S>();
  const S cutoff2 = A.cutoff2;

  for (int k = tid; k < E * E; k += blockDim.x) rv[k] = P.rvdw_e[k];

  for (;;) {
    __syncthreads();                       // previous centre completely finished (list, wiS, isum reusable)
    if (tid == 0) s_center = atomicAdd(Q.counter, 1);
    __syncthreads();
    const int c = s_center;
    if (c >= nrows) break;
    const int i = A.row_begin + c;
    const int ei = P.eidx[i];
    if (ei < 0) continue;                  // padding row (CTA-uniform)
    const Atom4<S> ai = A.atoms[i];
    const S gi = A.aux[i].b;
    if (tid < 8) { wiS[tid] = A.w[8 * (size_t)i + tid]; wiS[8 + tid] = A.dw[8 * (size_t)i + tid]; }
    const S ibox[6] = {ai.x, ai.y, ai.z, ai.x, ai.y, ai.z};

    // ---- neighbours of i within the cutoff, compacted in index order
    int n = 0;
    for (int js0 = 0; js0 < nsub; js0 += nw) {
      const int js = js0 + wid;
      unsigned m = 0;
      if (js < nsub && box_dist2(ibox, A.box + (size_t)js * 6) <= cutoff2) {
        const int j = js * kSysSub + lane;
        const Atom4<S> aj = A.atoms[j];
        const S dx = ai.x - aj.x, dy = ai.y - aj.y, dz = ai.z - aj.z;
        const bool ok = P.eidx[j] >= 0 && j != i && (dx * dx + dy * dy + dz * dz) <= cutoff2;
        m = __ballot_sync(0xffffffffu, ok);
      }
      if (lane == 0) cnt[wid] = __popc(m);
      __syncthreads();
      int base = n, tot = 0;
#pragma unroll
      for (int w2 = 0; w2 < nw; ++w2) { const int cw = cnt[w2]; base += w2 < wid ? cw : 0; tot += cw; }
      if (m & (1u << lane)) list[base + __popc(m & ((1u << lane) - 1))] = js * kSysSub + lane;
      n += tot;
      __syncthreads();
    }
    // ---- one record per neighbour (edge i-n constants)
    for (int p = tid; p < n; p += blockDim.x) {
      const int nn = list[p];
      const Atom4<S> an = A.atoms[nn];
      const S dx = ai.x - an.x, dy = ai.y - an.y, dz = ai.z - an.z;
      const S r2 = dx * dx + dy * dy + dz * dz + tiny;
      const int en = P.eidx[nn];
      const S* Tn = P.tvec + ((size_t)nn * E + ei) * 8;     // sum_b K[Zi,Zn,a,b] w_nb
      const S* Tdn = tdvec + ((size_t)nn * E + ei) * 8;     // ... with dw_n
      S c6 = S(0), d_i = S(0), d_n = S(0);
#pragma unroll
      for (int a = 0; a < kNRef; ++a) { c6 += wiS[a] * Tn[a]; d_i += wiS[8 + a] * Tn[a]; d_n += wiS[a] * Tdn[a]; }
      const S ic6 = frcp(c6);
      S* r = rec + (size_t)p * kAtm3Rec;
      r[0] = sqr(ab(c6) + tiny);
      r[1] = r2;
      r[2] = frcp(r2);
      r[3] = d_i * ic6;
      r[4] = d_n * ic6;
      r[5] = rv[ei * E + en];
    }
    if (tid == 0) s_tile = (n + 31) / 32 - 1;
    __syncthreads();

    // ---- tiles of 32 list entries (lane <-> k), largest tile first
    S Eia = S(0);                                      // per lane: sum e * 2 m_jk
    S Gix = S(0), Giy = S(0), Giz = S(0), Di = S(0);   // per lane (k side) + lane 0 (j side)
    for (;;) {
      int t = 0;
      if (lane == 0) t = atomicSub(&s_tile, 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t < 0) break;
      const int kpos = 32 * t + lane;
      const bool kvalid = kpos < n;
      const int kp = kvalid ? kpos : 32 * t;
      const int kidx = list[kp];
      const Atom4<S> ak = A.atoms[kidx];
      const int ek = P.eidx[kidx];
      const S gk = A.aux[kidx].b;
      S wk[kNRef], dwk[kNRef];
#pragma unroll
      for (int a = 0; a < kNRef; ++a) { wk[a] = A.w[8 * (size_t)kidx + a]; dwk[a] = WANT_G ? A.dw[8 * (size_t)kidx + a] : S(0); }
      const S* rk = rec + (size_t)kp * kAtm3Rec;
      const S hik = rk[0], b2 = rk[1], inv_b = rk[2], Rik = rk[5];
      S Sfb = S(0), Swe = S(0), Ek = S(0), Vx = S(0), Vy = S(0), Vz = S(0), Dk = S(0);
      const int jend = min(32 * t + 31, n - 1);        // positions j < k only
      for (int jpos = 0; jpos < jend; ++jpos) {
        const int jidx = list[jpos];                   // warp-uniform
        const Atom4<S> aj = A.atoms[jidx];
        const S dxjk = aj.x - ak.x, dyjk = aj.y - ak.y, dzjk = aj.z - ak.z;
        const S c2e = dxjk * dxjk + dyjk * dyjk + dzjk * dzjk;
        const bool mjk = c2e <= cutoff2;
        // a triangle belongs to its smallest index: i < j (< k, the list is sorted)
        const bool act = kvalid && kpos > jpos && (!mjk || jidx > i);
        if (!__any_sync(0xffffffffu, act)) continue;
        const S* rj = rec + (size_t)jpos * kAtm3Rec;
        const S hij = rj[0], a2 = rj[1], inv_a = rj[2];
        const int ej = P.eidx[jidx];
        const S gj = A.aux[jidx].b;
        S vals[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) vals[v] = S(0);
        if (act) {
          const Pair2<S>* Tj2 = reinterpret_cast<const Pair2<S>*>(P.tvec + ((size_t)jidx * E + ek) * 8);
          S Tj[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) { const Pair2<S> pp = Tj2[q]; Tj[2 * q] = pp.a; Tj[2 * q + 1] = pp.b; }
          S c6jk = wk[0] * Tj[0];
#pragma unroll
          for (int a = 1; a < kNRef; ++a) c6jk += wk[a] * Tj[a];
          const S xjk = ab(c6jk) + tiny;
          const S hjk = xjk * frsq(xjk);
          const S c9 = P.s9 * (hij * (hik * hjk));
          const S r0 = (rs93 * rj[5]) * (Rik * rv[ej * E + ek]);
          const S c2 = c2e + tiny;
          const S r2 = a2 * b2 * c2;
          const S r1inv = frsq(r2);
          const S r1inv2 = r1inv * r1inv;
          const S r3inv = r1inv2 * r1inv;
          const S r5inv = r3inv * r1inv2;
          const S p = a2 + c2 - b2, q = a2 - c2 + b2, tt = (c2 + b2) - a2;
          const S s = p * q * tt;
          const S sr5 = s * r5inv;
          const S ang = S(0.375) * sr5 + r3inv;
          const S base = r0 * r1inv;
          const S u = P.cube_root_pow ? pow16_3<S>(base) : fex(fmx(P.pexp * lg(base), S(-700)));
          const S fd = frcp(S(1) + S(6) * u);
          const S e = c9 * (ang * fd);
          const S share = mjk ? S(2) : S(1);           // 1 + m_jk
          const S es = e * share;
          Eia += mjk ? e + e : S(0);
          Ek += es;
          vals[0] = es;
          if (WANT_G) {
            const S dfd0 = (S(3) * P.pexp) * (u * (fd * fd));
            const S qt = q * tt, pt = p * tt, pq = p * q;
            const S k0 = c9 * fd, k1 = c9 * (ang * dfd0);
            const S common = k1 - k0 * (S(0.9375) * sr5 + S(1.5) * r3inv);
            const S k0r5 = S(0.375) * (k0 * r5inv);
            const S inv_c = r1inv2 * (a2 * b2);
            const S w6 = (mjk ? gi + gi : S(0)) + (gj + gk) * share;
            const S we = w6 * e;
            const S fa = w6 * (k0r5 * (qt + pt - pq) + common * inv_a);
            const S fb = w6 * (k0r5 * (pt + pq - qt) + common * inv_b);
            const S fc = w6 * (k0r5 * (qt + pq - pt) + common * inv_c);
            Sfb += fb;
            Swe += we;
            const S vx = fc * dxjk, vy = fc * dyjk, vz = fc * dzjk;
            Vx += vx; Vy += vy; Vz += vz;
            const Pair2<S>* Td2 = reinterpret_cast<const Pair2<S>*>(tdvec + ((size_t)jidx * E + ek) * 8);
            S Td[8];
#pragma unroll
            for (int q2 = 0; q2 < 4; ++q2) { const Pair2<S> pp = Td2[q2]; Td[2 * q2] = pp.a; Td[2 * q2 + 1] = pp.b; }
            S dk = dwk[0] * Tj[0], dj = wk[0] * Td[0];
#pragma unroll
            for (int a = 1; a < kNRef; ++a) { dk += dwk[a] * Tj[a]; dj += wk[a] * Td[a]; }
            const S qc = we * frcp(c6jk);
            Dk += qc * dk;
            vals[1 % NV] = fa; vals[2 % NV] = we; vals[3 % NV] = vx; vals[4 % NV] = vy; vals[5 % NV] = vz;
            vals[6 % NV] = qc * dj;
          }
        }
        // ---- j side of this step: sums over the 32 lanes, five atomics
        const S tot = warp_rows_reduce<NV, S>(vals, redw, lane);
        const S EJ = bcast(tot, 0);
        S out = sixth * EJ;
        S* addr = A.energy + jidx;
        if (WANT_G) {
          const S FA = bcast(tot, 4), WE = bcast(tot, 8), VX = bcast(tot, 12), VY = bcast(tot, 16),
                  VZ = bcast(tot, 20), DJ = bcast(tot, 24);
          const S dxij = ai.x - aj.x, dyij = ai.y - aj.y, dzij = ai.z - aj.z;
          if (lane == 0) { Gix += FA * dxij; Giy += FA * dyij; Giz += FA * dzij; Di += WE * rj[3]; }
          if (lane == 1) { out = two6 * (VX - FA * dxij); addr = A.grad + 3 * (size_t)jidx; }
          if (lane == 2) { out = two6 * (VY - FA * dyij); addr = A.grad + 3 * (size_t)jidx + 1; }
          if (lane == 3) { out = two6 * (VZ - FA * dzij); addr = A.grad + 3 * (size_t)jidx + 2; }
          if (lane == 4) { out = half6 * (WE * rj[4] + DJ); addr = A.decn + jidx; }
        }
        if (lane < (WANT_G ? 5 : 1)) red_add(addr, out);
      }
      // ---- k side of the tile
      if (kvalid) {
        red_add(A.energy + kidx, sixth * Ek);
        if (WANT_G) {
          const S dxik = ai.x - ak.x, dyik = ai.y - ak.y, dzik = ai.z - ak.z;
          red_add(A.grad + 3 * (size_t)kidx, -two6 * (Sfb * dxik + Vx));
          red_add(A.grad + 3 * (size_t)kidx + 1, -two6 * (Sfb * dyik + Vy));
          red_add(A.grad + 3 * (size_t)kidx + 2, -two6 * (Sfb * dzik + Vz));
          red_add(A.decn + kidx, half6 * (Swe * rk[4] + Dk));
          Gix += Sfb * dxik; Giy += Sfb * dyik; Giz += Sfb * dzik;
          Di += Swe * rk[3];
        }
      }
    }
    // ---- i side of the centre
    {
      S v[5] = {Eia, Gix, Giy, Giz, Di};
#pragma unroll
      for (int q = 0; q < (WANT_G ? 5 : 1); ++q) {
        const S s = warp_sum<S>(v[q]);
        if (lane == 0) isum[wid][q] = s;
      }
      __syncthreads();
      if (tid < (WANT_G ? 5 : 1)) {
        S s = S(0);
#pragma unroll
        for (int w2 = 0; w2 < nw; ++w2) s += isum[w2][tid];
        if (tid == 0) red_add(A.energy + i, sixth * s);
        else if (tid < 4) red_add(A.grad + 3 * (size_t)i + (tid - 1), two6 * s);
        else red_add(A.decn + i, half6 * s);
      }
    }
  }
}

}  // namespace d3
