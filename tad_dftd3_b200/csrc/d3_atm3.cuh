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
//    centres of the 32-atom sub-tiles row_begin/32 + k * tile_stride below row_end, which
//    is also how ranks shard the term: results are ADDED, so an all-reduce follows);
//  * the CTA compacts the neighbours of i (r <= c), in index order, into its own
//    list in global scratch, with one self-contained record per neighbour n:
//    {x, y, z, (index, element)}, {sqrt|C6_in|, r_in^2, 1/r_in^2, (dC6_in/dCN_i)/C6_in,
//    (dC6_in/dCN_n)/C6_in, R_in, g_n}, so the sweeps read one contiguous stream;
//  * warps take tiles of 64 list entries (largest first): each lane owns two of them
//    (two independent evaluation chains per step) and walks all list positions j
//    before its k (j is warp-uniform).  Everything that belongs
//    to k or to edge ik accumulates in registers over the j loop; the j-side sums
//    of a step (7 values) go through a transposed shared-memory reduction and are
//    added to the global arrays with five fp64 atomics per 32 triples; the i-side
//    sums stay in registers until the centre is finished;
//  * C6_jk and its two CN derivatives are 7-wide dots of the lane's w_k / dw_k
//    with the pre-contracted vectors T_j[e_k], Td_j[e_k] (`system_tvec_kernel`);
//  * TAB variant: everything of a triple that depends on ONE edge is the same for every centre whose list holds
//    both ends (hundreds at a 25 Bohr cutoff): sqrt|C6_jk| and the two logarithmic CN derivatives of C6_jk, and,
//    because the damping argument and the distance product factorise over the edges,
//      (r0/r)^p = prod_ab (rs9 R_ab / r_ab)^p,
//    also U_ab = (rs9 R_ab / r_ab)^p.  `system_atm_pairtab_kernel` computes them ONCE per pair into a dense
//    [natoms][natoms] table of 32-byte entries (upper triangle); the records carry U for the edges at the
//    centre.  The sweep then reads one sector per lane and triple instead of three 7-wide dots, a square
//    root, a reciprocal and the power, and needs the values only at the end of a triple's dependency chain.
//    Used while the table fits `atm_work` (d3b200.h).
#pragma once
#include "d3_atm.cuh"

namespace d3 {

#ifndef D3_ATM3_WARPS
#define D3_ATM3_WARPS 4
#define D3_ATM3_CTAS 3
#endif
#ifndef D3_ATM3_KPL
#define D3_ATM3_KPL 2
#endif
constexpr int kAtm3Kpl = D3_ATM3_KPL;
constexpr int kAtm3Warps = D3_ATM3_WARPS;
constexpr int kAtm3Ctas = D3_ATM3_CTAS;         // resident CTAs per SM the kernel is built for
#ifndef D3_ATM3_TAB_CTAS
#define D3_ATM3_TAB_CTAS D3_ATM3_CTAS
#endif
constexpr int kAtm3TabCtas = D3_ATM3_TAB_CTAS;  // ... of the pair-table variant (fewer live registers)
#ifndef D3_ATM3_PF
#define D3_ATM3_PF 1                            // > 0: L2 prefetch of the table entries of the step this many ahead
#endif
#ifndef D3_ATM3_JEARLY
#define D3_ATM3_JEARLY 1                        // whole record of list entry j loaded at the top of a step
#endif



constexpr int kAtm3Rec = 12;         // reals per neighbour record (11 used), 16-byte aligned pairs
constexpr int kAtm3RedStride = 40;   // row stride of the reduction tile (rows 16 banks apart)
constexpr int kAtm3Tab = 4;          // reals per entry of the pair table (one 32-byte sector in fp64)

template <typename S>
struct AtmArgs3 {
  AtmArgs2<S> a;
  int* list;       // [gridDim.x][natoms]          neighbour indices of the CTA's current centre
  S* rec;          // [gridDim.x][natoms][kAtm3Rec] neighbour records
  int* counter;    // zero before the launch; counter[1] = largest neighbour count of any atom (system_atm_maxn_kernel)
  int tile_stride; // > 0: the launch owns the 32-atom sub-tiles row_begin/32 + k * tile_stride below row_end;
                   // < 0: the atoms row_begin + k * |tile_stride| below row_end
  int parts;       // a centre is split into `parts` work units (every parts-th tile of its list): keeps
                   // the persistent CTAs busy when a launch owns few centres (several GPUs)
  const S* tab;    // TAB variant: [natoms][natoms][kAtm3Tab] pair table of `system_atm_pairtab_kernel`, else unused
};

template <typename S>
inline size_t atm3_workspace_bytes(int natoms, int grid) {
  const size_t b = 256 + (size_t)grid * natoms * (sizeof(int) + kAtm3Rec * sizeof(S));
  return (b + 255) / 256 * 256;
}
template <typename S>
inline size_t atm3_pairtab_bytes(int natoms) { return (size_t)natoms * natoms * kAtm3Tab * sizeof(S); }

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

// One 16-byte half of a pair-table entry, loaded only where `on` (zeros elsewhere), around L1 (ld.global.cg: an
// entry is read once per centre, and L1 is what keeps the CTA's neighbour records).
__device__ __forceinline__ Pair2<double> tab_load(const Pair2<double>* p, bool on) {
  Pair2<double> r;
  r.a = r.b = 0.0;
  if (on) { const double2 v = __ldcg(reinterpret_cast<const double2*>(p)); r.a = v.x; r.b = v.y; }
  return r;
}
__device__ __forceinline__ Pair2<float> tab_load(const Pair2<float>* p, bool on) {
  Pair2<float> r;
  r.a = r.b = 0.0f;
  if (on) { const float2 v = __ldcg(reinterpret_cast<const float2*>(p)); r.a = v.x; r.b = v.y; }
  return r;
}

__device__ __forceinline__ float bcast(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// bit casts for the (index, element) slots of a record (moved, never computed with)
__device__ __forceinline__ float int_as(int v, float) { return __int_as_float(v); }
__device__ __forceinline__ double int_as(int v, double) { return __hiloint2double(0, v); }
__device__ __forceinline__ int as_int(float v) { return __float_as_int(v); }
__device__ __forceinline__ int as_int(double v) { return __double2loint(v); }

// x^(16/3), x > 0: x^(-1/3) from two MUFU approximations, one cubic step, then (x y^2)^16
template <typename S>
__device__ __forceinline__ S pow16_3_fast(S x) {
  float l, yf;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"((float)x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"(-0.33333334f * l));
  S y = S(yf);
  const S y3 = y * y * y;
  const S e = S(1) - x * y3;
  y = y + (y * e) * (S(1.0 / 3.0) + S(2.0 / 9.0) * e);
  const S c = x * (y * y);          // x^(1/3)
  const S c2 = c * c, c4 = c2 * c2, c8 = c4 * c4;
  return c8 * c8;
}

// (rs9 R / r)^p of one edge (damping/atm.py:133-135 per edge); the power of a triple is the product of its three
template <typename S>
__device__ __forceinline__ S atm_edge_pow(S base, S pexp, int cube_root_pow) {
  return cube_root_pow ? pow16_3_fast<S>(base) : fex(fmx(pexp * lg(base), S(-700)));
}

// Largest number of neighbours within the cutoff of any atom: sizes the per-CTA slots of the neighbour scratch
// of `system_atm3_kernel` (slots of `natoms` entries, 1 MB apart, alias in the L2 sets and were written back
// to DRAM at 50x the algorithmic bytes, profiles/r1h_system_atm3_raw.csv).  One warp per 32-atom sub-tile,
// lane = atom, partner sub-tiles culled by their boxes; the result is an atomicMax into *out (zero before).
constexpr int kMaxnWarps = 8;
template <typename S>
__global__ void __launch_bounds__(32 * kMaxnWarps) system_atm_maxn_kernel(SysArgs<S> A, const signed char* eidx, int* out) {
  // one CTA per 32-atom sub-tile (lane = atom), its warps deal the partner sub-tiles round-robin
  __shared__ int cnt[kMaxnWarps][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int sub = blockIdx.x;
  const int nsub = A.natoms / kSysSub;
  const int i = sub * kSysSub + lane;
  const Atom4<S> ai = A.atoms[i];
  const S* ibox = A.box + (size_t)sub * 6;
  int n = 0;
  for (int js = w; js < nsub; js += kMaxnWarps) {
    if (box_dist2(ibox, A.box + (size_t)js * 6) > A.cutoff2) continue;     // warp-uniform
    const int j = js * kSysSub + lane;
    const Atom4<S> aj = A.atoms[j];
    const bool real_j = eidx[j] >= 0;
    for (int l = 0; l < 32; ++l) {
      const S x = __shfl_sync(0xffffffffu, aj.x, l), y = __shfl_sync(0xffffffffu, aj.y, l), z = __shfl_sync(0xffffffffu, aj.z, l);
      const bool rj = __shfl_sync(0xffffffffu, (int)real_j, l) != 0;
      const S dx = ai.x - x, dy = ai.y - y, dz = ai.z - z;
      n += (rj && (js * kSysSub + l) != i && dx * dx + dy * dy + dz * dz <= A.cutoff2) ? 1 : 0;
    }
  }
  cnt[w][lane] = n;
  __syncthreads();
  if (w == 0) {
    int tot = 0;
#pragma unroll
    for (int k = 0; k < kMaxnWarps; ++k) tot += cnt[k][lane];
    tot = eidx[i] >= 0 ? tot : 0;
    tot = __reduce_max_sync(0xffffffffu, tot);
    if (lane == 0 && tot > 0) atomicMax(out, tot);
  }
}

// Pair table of the TAB variant: entry [j][k] (k > j; the blocks on the diagonal are written completely) =
//   { sqrt(|C6_jk| + tiny), U_jk = (rs9 R_jk / r_jk)^p, (dC6_jk/dCN_j) / C6_jk, (dC6_jk/dCN_k) / C6_jk }
// with C6_jk = sum_a w_k[a] T_j[e_k][a] in the summation order of the sweep (damping/atm.py:93-100, model/c6.py:81-87).
// One CTA per (128 columns k, 32 rows j); a thread owns a column, the T / Td vectors of the rows sit in shared memory.
constexpr int kPairTabCols = 128, kPairTabRows = 32;
template <typename S>
__global__ void __launch_bounds__(kPairTabCols) system_atm_pairtab_kernel(AtmArgs2<S> P, S* tab) {
  __shared__ __align__(16) S sT[kPairTabRows * kAtmEmax * 8];
  __shared__ __align__(16) S sTd[kPairTabRows * kAtmEmax * 8];
  __shared__ S rv[kAtmEmax * kAtmEmax];
  __shared__ int sej[kPairTabRows];
  __shared__ S sx[kPairTabRows], sy[kPairTabRows], sz[kPairTabRows];
  const SysArgs<S>& A = P.sys;
  const int natoms = A.natoms, E = P.nelem;
  const int k0 = blockIdx.x * kPairTabCols, j0 = blockIdx.y * kPairTabRows;
  if (k0 + kPairTabCols <= j0) return;                  // block entirely below the diagonal
  const int tid = threadIdx.x;
  const size_t toff = (size_t)j0 * E * 8, tdoff = (size_t)natoms * E * 8;
  for (int q = tid; q < kPairTabRows * E * 8; q += kPairTabCols) { sT[q] = P.tvec[toff + q]; sTd[q] = P.tvec[tdoff + toff + q]; }
  for (int q = tid; q < E * E; q += kPairTabCols) rv[q] = P.rvdw_e[q];
  if (tid < kPairTabRows) {
    sej[tid] = P.eidx[j0 + tid];
    const Atom4<S> aj = A.atoms[j0 + tid];
    sx[tid] = aj.x; sy[tid] = aj.y; sz[tid] = aj.z;
  }
  __syncthreads();
  const int k = k0 + tid;
  if (k >= natoms) return;
  const int ek = P.eidx[k];
  if (ek < 0) return;                                   // padding column: never in a neighbour list
  S wk[kNRef], dwk[kNRef];
#pragma unroll
  for (int a = 0; a < kNRef; ++a) { wk[a] = A.w[8 * (size_t)k + a]; dwk[a] = A.dw[8 * (size_t)k + a]; }
  const Atom4<S> ak = A.atoms[k];
  const S tiny = tiny_of<S>();
  for (int jj = 0; jj < kPairTabRows; ++jj) {
    const int ej = sej[jj];
    if (ej < 0) continue;                               // CTA-uniform
    const S* T = sT + (jj * E + ek) * 8;
    const S* Td = sTd + (jj * E + ek) * 8;
    S c6 = wk[0] * T[0], dk = dwk[0] * T[0], dj = wk[0] * Td[0];
#pragma unroll
    for (int a = 1; a < kNRef; ++a) { c6 += wk[a] * T[a]; dk += dwk[a] * T[a]; dj += wk[a] * Td[a]; }
    const S x = ab(c6) + tiny;
    const S ic = rcp0(c6);
    const S dx = sx[jj] - ak.x, dy = sy[jj] - ak.y, dz = sz[jj] - ak.z;
    const S rinv = frsq(dx * dx + dy * dy + dz * dz + tiny);
    const S U = atm_edge_pow<S>((P.rs9 * rv[ej * E + ek]) * rinv, P.pexp, P.cube_root_pow);
    Pair2<S>* out = reinterpret_cast<Pair2<S>*>(tab + ((size_t)(j0 + jj) * natoms + k) * kAtm3Tab);
    Pair2<S> q;
    q.a = x * frsq(x); q.b = U; out[0] = q;
    q.a = dj * ic; q.b = dk * ic; out[1] = q;
  }
}

// record slots
enum { kRx = 0, kRy, kRz, kRidx, kRh, kRr2, kRinv, kRci, kRcn, kRvdw, kRg, kRe };

template <typename S, bool WANT_G, bool TAB = false>
__global__ void __launch_bounds__(32 * kAtm3Warps, TAB ? kAtm3TabCtas : kAtm3Ctas) system_atm3_kernel(AtmArgs3<S> Q) {
  const AtmArgs2<S>& P = Q.a;
  const SysArgs<S>& A = P.sys;
  constexpr int nw = kAtm3Warps;
  constexpr int NV = WANT_G ? 7 : 1;
  constexpr int KPL = kAtm3Kpl;                      // list entries per lane
  constexpr int TK = 32 * KPL;
  __shared__ S rv[kAtmEmax * kAtmEmax];
  __shared__ __align__(16) S wiS[16];
  __shared__ __align__(16) S red[nw * 7 * kAtm3RedStride];
  __shared__ S wsh[TAB ? 1 : nw][KPL][2 * kNRef][32];   // w_k, dw_k of each warp's tile, [k of the lane][a][lane]
  __shared__ S isum[nw][5];
  __shared__ int cnt[nw];
  __shared__ int s_center, s_tile;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int E = P.nelem;
  const int natoms = A.natoms;
  const int nsub = natoms / kSysSub;
  // ownership of the centres: tile_stride > 0: the 32-atom sub-tiles row_begin/32 + k * tile_stride;
  // tile_stride < 0: the atoms row_begin + k * |tile_stride| (atom by atom: neighbouring atoms of the sorted
  // order cost about the same, so ranks that deal them round-robin finish together)
  const bool atomwise = Q.tile_stride < 0;
  const int tstride = atomwise ? -Q.tile_stride : (Q.tile_stride > 0 ? Q.tile_stride : 1);
  const int nsub_own = ((A.row_end - A.row_begin) / kSysSub + tstride - 1) / tstride;
  const int nrows = atomwise ? (A.row_end - A.row_begin + tstride - 1) / tstride : nsub_own * kSysSub;
  // per-CTA slots sized by the largest neighbour count (rounded up to whole tiles), packed back to back
  const int slot = min(natoms, ((Q.counter[1] + TK - 1) / TK) * TK);
  int* list = Q.list + (size_t)blockIdx.x * slot;
  S* rec = Q.rec + (size_t)blockIdx.x * slot * kAtm3Rec;
  S* redw = red + wid * 7 * kAtm3RedStride;
  const S* __restrict__ tvec = P.tvec;
  const unsigned tdoff = (unsigned)natoms * E * 8;  // Td vectors sit right behind the T vectors
  const S sixth = S(1) / S(6), two6 = S(2) * sixth, half6 = S(0.5) * sixth;
  const S rs93 = P.rs9 * P.rs9 * P.rs9;
  const S tiny = tiny_of<S>();
  const S cutoff2 = A.cutoff2;

  for (int k = tid; k < E * E; k += blockDim.x) rv[k] = P.rvdw_e[k];
  // j side of a step: what lane 0..4 adds (row of the reduced sums, factor, target, stride)
  const int jsrc = lane == 0 ? 0 : (lane == 4 ? 24 : 4 * (2 + (lane & 3)));
  const S jk1 = lane == 0 ? sixth : (lane == 4 ? half6 : two6);
  S* const jptr = lane == 0 ? A.energy : (lane == 4 ? A.decn : A.grad + ((lane - 1) & 3));
  const int jmul = (lane >= 1 && lane <= 3) ? 3 : 1;

  for (;;) {
    __syncthreads();                       // previous centre completely finished (list, wiS, isum reusable)
    if (tid == 0) s_center = atomicAdd(Q.counter, 1);
    __syncthreads();
    const int parts = Q.parts > 0 ? Q.parts : 1;
    const int part = s_center % parts;
    const int c = s_center / parts;
    if (c >= nrows) break;
    const int i = atomwise ? A.row_begin + c * tstride : A.row_begin + (c / kSysSub) * tstride * kSysSub + (c % kSysSub);
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
    // ---- one record per neighbour: its own data and the constants of edge i-n
    for (int p = tid; p < n; p += blockDim.x) {
      const int nn = list[p];
      const Atom4<S> an = A.atoms[nn];
      const S dx = ai.x - an.x, dy = ai.y - an.y, dz = ai.z - an.z;
      const S r2 = dx * dx + dy * dy + dz * dz + tiny;
      const int en = P.eidx[nn];
      const S* Tn = tvec + ((size_t)nn * E + ei) * 8;      // sum_b K[Zi,Zn,a,b] w_nb
      const S* Tdn = Tn + tdoff;                           // ... with dw_n
      S c6 = S(0), d_i = S(0), d_n = S(0);
#pragma unroll
      for (int a = 0; a < kNRef; ++a) { c6 += wiS[a] * Tn[a]; d_i += wiS[8 + a] * Tn[a]; d_n += wiS[a] * Tdn[a]; }
      const S ic6 = rcp0(c6);
      Pair2<S>* r = reinterpret_cast<Pair2<S>*>(rec + (size_t)p * kAtm3Rec);
      Pair2<S> q;
      q.a = an.x; q.b = an.y; r[0] = q;
      q.a = an.z; q.b = int_as(nn, S(0)); r[1] = q;
      q.a = sqr(ab(c6) + tiny); q.b = r2; r[2] = q;
      q.a = frcp(r2); q.b = d_i * ic6; r[3] = q;
      if (TAB) {
        // edge factor of the damping power
        q.a = d_n * ic6; q.b = atm_edge_pow<S>((P.rs9 * rv[ei * E + en]) * frsq(r2), P.pexp, P.cube_root_pow); r[4] = q;
        q.a = A.aux[nn].b; q.b = S(0); r[5] = q;
      } else {
        q.a = d_n * ic6; q.b = rv[ei * E + en]; r[4] = q;
        q.a = A.aux[nn].b; q.b = int_as(en, S(0)); r[5] = q;
      }
    }
    const int ntile = (n + TK - 1) / TK;
    if (tid == 0) s_tile = (ntile - 1 - part + parts) / parts - 1;   // tiles part, part + parts, ... of this unit
    __syncthreads();

    // ---- tiles of 32 * KPL list entries (each lane owns KPL of them), largest tile first
    S Eia = S(0);                                      // per lane: sum e * 2 m_jk
    S Gix = S(0), Giy = S(0), Giz = S(0), Di = S(0);   // per lane (k side) + lane 0 (j side)
    for (;;) {
      int t = 0;
      if (lane == 0) t = atomicSub(&s_tile, 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t < 0) break;
      t = t * parts + part;
      int kpos[KPL];
      bool kvalid[KPL];
      const Pair2<S>* rk[KPL];
      S akx[KPL], aky[KPL], akz[KPL], hik[KPL], b2[KPL], inv_b[KPL], Rik[KPL], gk[KPL];   // Rik: TAB: U_ik
      unsigned ekoff[KPL];       // !TAB: offset of T_j[e_k]; TAB: offset of column k in a table row
      const S* rvk[KPL];
      S Sfb[KPL], Swe[KPL], Ek[KPL], Vx[KPL], Vy[KPL], Vz[KPL], Dk[KPL];
      __syncwarp();
#pragma unroll
      for (int kk = 0; kk < KPL; ++kk) {
        kpos[kk] = TK * t + 32 * kk + lane;
        kvalid[kk] = kpos[kk] < n;
        rk[kk] = reinterpret_cast<const Pair2<S>*>(rec + (size_t)(kvalid[kk] ? kpos[kk] : TK * t) * kAtm3Rec);
        const Pair2<S> k0_ = rk[kk][0], k2_ = rk[kk][2];
        akx[kk] = k0_.a; aky[kk] = k0_.b; akz[kk] = rk[kk][1].a;
        hik[kk] = k2_.a; b2[kk] = k2_.b; inv_b[kk] = rk[kk][3].a; Rik[kk] = rk[kk][4].b; gk[kk] = rk[kk][5].a;
        const int kidx = as_int(rk[kk][1].b), ek = TAB ? 0 : as_int(rk[kk][5].b);
        if (!TAB) {
#pragma unroll
          for (int a = 0; a < kNRef; ++a) {
            wsh[wid][kk][a][lane] = A.w[8 * (size_t)kidx + a];
            if (WANT_G) wsh[wid][kk][kNRef + a][lane] = A.dw[8 * (size_t)kidx + a];
          }
        }
        ekoff[kk] = TAB ? (unsigned)kidx * (unsigned)kAtm3Tab : (unsigned)ek * 8;
        rvk[kk] = rv + ek;
        Sfb[kk] = Swe[kk] = Ek[kk] = Vx[kk] = Vy[kk] = Vz[kk] = Dk[kk] = S(0);
      }
      __syncwarp();
      const int jend = min(TK * t + TK - 1, n - 1);    // positions j < k only
      const Pair2<S>* rj = reinterpret_cast<const Pair2<S>*>(rec);
      for (int jpos = 0; jpos < jend; ++jpos, rj += kAtm3Rec / 2) {
        const Pair2<S> j0_ = rj[0], j1_ = rj[1];         // warp-uniform loads
#if D3_ATM3_JEARLY
        // the rest of the record with them (skipped steps are rare), so that its latency overlaps the distances
        const Pair2<S> j2_ = rj[2], j3_ = rj[3], j4_ = rj[4], j5_ = rj[5];
#endif
        const int jidx = as_int(j1_.b);
        S dxjk[KPL], dyjk[KPL], dzjk[KPL], c2e[KPL];
        bool mjk[KPL], act[KPL];
        bool any_act = false;
#pragma unroll
        for (int kk = 0; kk < KPL; ++kk) {
          dxjk[kk] = j0_.a - akx[kk]; dyjk[kk] = j0_.b - aky[kk]; dzjk[kk] = j1_.a - akz[kk];
          c2e[kk] = dxjk[kk] * dxjk[kk] + dyjk[kk] * dyjk[kk] + dzjk[kk] * dzjk[kk];
          mjk[kk] = c2e[kk] <= cutoff2;
          // a triangle belongs to its smallest index: i < j (< k, the list is sorted)
          act[kk] = kvalid[kk] && kpos[kk] > jpos && (!mjk[kk] || jidx > i);
          any_act = any_act || act[kk];
        }
        // TAB: {sqrt|C6_jk|, U_jk}, {dlnC6/dCN_j, dlnC6/dCN_k} of the counted triples (zeros elsewhere), requested
        // at the top of the step, before the vote: a whole dependency chain ahead of their first use (inside the
        // per-triple code the scheduler sank them to ~90 instructions before the use: 43.5 -> 39.0 ms)
        Pair2<S> tb0[KPL], tb1[KPL];
        if (TAB) {
          const S* tabrow = Q.tab + (size_t)jidx * natoms * kAtm3Tab;
#pragma unroll
          for (int kk = 0; kk < KPL; ++kk) {
            const Pair2<S>* tp = reinterpret_cast<const Pair2<S>*>(tabrow + ekoff[kk]);
            tb0[kk] = tab_load(tp, act[kk]);
            if (WANT_G) tb1[kk] = tab_load(tp + 1, act[kk]);
          }
        }
#if D3_ATM3_PF
        // a quarter of the entries comes from DRAM (L2 hit rate 74 %, profiles/r2_atm3_tab_raw.csv): ask L2 for the
        // entries of a coming step (all lanes behind it, counted or not); 39.0 -> 37.4 ms
        if (TAB && jpos + D3_ATM3_PF < jend) {
          const int jn = as_int(rj[D3_ATM3_PF * (kAtm3Rec / 2) + 1].b);
          const S* rown = Q.tab + (size_t)jn * natoms * kAtm3Tab;
#pragma unroll
          for (int kk = 0; kk < KPL; ++kk)
            if (kvalid[kk] && kpos[kk] > jpos + D3_ATM3_PF) asm volatile("prefetch.global.L2 [%0];" ::"l"(rown + ekoff[kk]));
        }
#endif
        if (!__any_sync(0xffffffffu, any_act)) continue;
#if !D3_ATM3_JEARLY
        const Pair2<S> j2_ = rj[2], j3_ = rj[3], j4_ = rj[4], j5_ = rj[5];
#endif
        const S hij = j2_.a, a2 = j2_.b, inv_a = j3_.a;
        const int ej = TAB ? 0 : as_int(j5_.b);
        const unsigned jbase = (unsigned)jidx * (unsigned)E * 8u;
        const S r0j = rs93 * j4_.b;                      // TAB: j4_.b = U_ij
        S vals[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) vals[v] = S(0);
        // the KPL triples of a lane are evaluated unconditionally (independent chains for the
        // scheduler); inactive ones are discarded by selects at the end
#pragma unroll
        for (int kk = 0; kk < KPL; ++kk) {
          const S* Tjp = tvec + (jbase + ekoff[kk]);
          S Tj[8];
          S c6jk = S(0), hjk;
          Pair2<S> t0, t1;
          if (TAB) {
            t0 = tb0[kk];
            if (WANT_G) t1 = tb1[kk];
            hjk = t0.a;
          } else {
            const Pair2<S>* Tj2 = reinterpret_cast<const Pair2<S>*>(Tjp);
#pragma unroll
            for (int q = 0; q < 4; ++q) { const Pair2<S> pp = Tj2[q]; Tj[2 * q] = pp.a; Tj[2 * q + 1] = pp.b; }
            c6jk = wsh[wid][kk][0][lane] * Tj[0];
#pragma unroll
            for (int a = 1; a < kNRef; ++a) c6jk += wsh[wid][kk][a][lane] * Tj[a];
            const S xjk = ab(c6jk) + tiny;
            hjk = xjk * frsq(xjk);
          }
          const S c9 = P.s9 * (hij * (hik[kk] * hjk));
          const S c2 = c2e[kk] + tiny;
          const S ab2 = a2 * b2[kk];
          const S r1inv = frsq(ab2 * c2);
          S u;
          if (TAB) {
            // inactive lanes: zeros from the table -> u = 0, fd = 1, everything selected away below
            u = (j4_.b * Rik[kk]) * t0.b;
          } else {
            const S r0 = r0j * (Rik[kk] * rvk[kk][ej * E]);
            u = atm_edge_pow<S>(r0 * r1inv, P.pexp, P.cube_root_pow);
          }
          const S r1inv2 = r1inv * r1inv;
          const S r3inv = r1inv2 * r1inv;
          const S r5inv = r3inv * r1inv2;
          const S p = a2 + c2 - b2[kk], q = a2 - c2 + b2[kk], tt = (c2 + b2[kk]) - a2;
          const S s = p * q * tt;
          const S sr5 = s * r5inv;
          const S ang = S(0.375) * sr5 + r3inv;
          const S fd = frcp(S(1) + S(6) * u);
          const S e = act[kk] ? c9 * (ang * fd) : S(0);
          const S share = mjk[kk] ? S(2) : S(1);       // 1 + m_jk
          const S es = e * share;
          Eia += mjk[kk] ? e + e : S(0);
          Ek[kk] += es;
          vals[0] += es;
          if (WANT_G) {
            const S dfd0 = (S(3) * P.pexp) * (u * (fd * fd));
            const S qt = q * tt, pt = p * tt, pq = p * q;
            const S k0 = c9 * fd, k1 = c9 * (ang * dfd0);
            const S common = k1 - k0 * (S(0.9375) * sr5 + S(1.5) * r3inv);
            const S k0r5 = S(0.375) * (k0 * r5inv);
            const S inv_c = r1inv2 * ab2;
            const S w6 = act[kk] ? (mjk[kk] ? gi + gi : S(0)) + (j5_.a + gk[kk]) * share : S(0);
            const S we = w6 * e;
            const S fa = act[kk] ? w6 * (k0r5 * (qt + pt - pq) + common * inv_a) : S(0);
            const S fb = act[kk] ? w6 * (k0r5 * (pt + pq - qt) + common * inv_b[kk]) : S(0);
            const S fc = act[kk] ? w6 * (k0r5 * (qt + pq - pt) + common * inv_c) : S(0);
            Sfb[kk] += fb;
            Swe[kk] += we;
            const S vx = fc * dxjk[kk], vy = fc * dyjk[kk], vz = fc * dzjk[kk];
            Vx[kk] += vx; Vy[kk] += vy; Vz[kk] += vz;
            vals[1 % NV] += fa; vals[2 % NV] += we; vals[3 % NV] += vx; vals[4 % NV] += vy; vals[5 % NV] += vz;
            if (TAB) {
              // inactive lanes: we == 0 and the table values are zeros (not loaded)
              Dk[kk] += we * t1.b;
              vals[6 % NV] += we * (t1.a + j4_.a);       // + dlnC6_ij/dCN_j: the whole dE/dCN_j of the step
            } else {
              const Pair2<S>* Td2 = reinterpret_cast<const Pair2<S>*>(Tjp + tdoff);
              S Td[8];
#pragma unroll
              for (int q2 = 0; q2 < 4; ++q2) { const Pair2<S> pp = Td2[q2]; Td[2 * q2] = pp.a; Td[2 * q2 + 1] = pp.b; }
              S dk = wsh[wid][kk][kNRef][lane] * Tj[0], dj = wsh[wid][kk][0][lane] * Td[0];
#pragma unroll
              for (int a = 1; a < kNRef; ++a) { dk += wsh[wid][kk][kNRef + a][lane] * Tj[a]; dj += wsh[wid][kk][a][lane] * Td[a]; }
              const S qc = act[kk] ? we * rcp0(c6jk) : S(0);
              Dk[kk] += qc * dk;
              vals[6 % NV] += qc * dj + we * j4_.a;
            }
          }
        }
        // ---- j side of this step: sums over the 32 lanes, five atomics
        const S tot = warp_rows_reduce<NV, S>(vals, redw, lane);
        if (WANT_G) {
          // lanes 0..4 add one sum each -- E_j, dE/dx_j, dE/dy_j, dE/dz_j, dE/dCN_j -- through ONE branch-free
          // formula k (X - FA dc): source row, factor, target array and stride are per-lane constants
          // (ternaries over the lane index compiled to ten branches per step: 0.62 stall cycles per issue on
          // branch resolving, profiles/r2_atm3_tab_raw.csv)
          const S FA = bcast(tot, 4), WE = bcast(tot, 8);
          const S X = bcast(tot, jsrc);
          const S dxij = ai.x - j0_.a, dyij = ai.y - j0_.b, dzij = ai.z - j1_.a;
          const S dc = lane == 1 ? dxij : (lane == 2 ? dyij : (lane == 3 ? dzij : S(0)));
          const S out = jk1 * (X - FA * dc);
          if (lane == 0) { Gix += FA * dxij; Giy += FA * dyij; Giz += FA * dzij; Di += WE * j3_.b; }
          if (lane < 5) red_add(jptr + (size_t)jidx * jmul, out);
        } else {
          const S EJ = bcast(tot, 0);
          if (lane == 0) red_add(A.energy + jidx, sixth * EJ);
        }
      }
      // ---- k side of the tile
#pragma unroll
      for (int kk = 0; kk < KPL; ++kk) {
        if (kvalid[kk]) {
          const int kidx = as_int(rk[kk][1].b);
          red_add(A.energy + kidx, sixth * Ek[kk]);
          if (WANT_G) {
            const S cki = rk[kk][3].b, ckk = rk[kk][4].a;
            const S dxik = ai.x - akx[kk], dyik = ai.y - aky[kk], dzik = ai.z - akz[kk];
            red_add(A.grad + 3 * (size_t)kidx, -two6 * (Sfb[kk] * dxik + Vx[kk]));
            red_add(A.grad + 3 * (size_t)kidx + 1, -two6 * (Sfb[kk] * dyik + Vy[kk]));
            red_add(A.grad + 3 * (size_t)kidx + 2, -two6 * (Sfb[kk] * dzik + Vz[kk]));
            red_add(A.decn + kidx, half6 * (Swe[kk] * ckk + Dk[kk]));
            Gix += Sfb[kk] * dxik; Giy += Sfb[kk] * dyik; Giz += Sfb[kk] * dzik;
            Di += Swe[kk] * cki;
          }
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
