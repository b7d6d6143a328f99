// Fused per-structure kernel, second generation (the bench / production path
// for energies + first derivatives in float/double, two-body + CN).
//
// Same phases as `molecule_kernel` (d3_molecule.cuh) but every unordered pair
// is evaluated ONCE (Newton's third law) instead of from both ends:
//
//  * rows are processed in blocks of 16; a warp's two half-warps own the same
//    16 rows i and split the partner range j > i by parity, so j is uniform per
//    half-warp (shared-memory reads stay broadcasts) and the triangular waste
//    is that of 16-row blocks (~85 % lane efficiency at N = 85);
//  * the i-side contributions stay in registers; the j-side contributions of
//    the 16 lanes are summed with a "transposed" shuffle reduction (V values
//    in V-1+... shuffle-adds instead of 4V) and added by single lanes to
//    per-warp private accumulators in shared memory: no atomics, fixed
//    summation order, bitwise reproducible;
//  * warps take row-block pairs (b, nb-1-b), which all carry the same work;
//  * rsqrt / reciprocal / exp are the branch-free MUFU-seeded versions of
//    d3_math.cuh, the two BJ reciprocals share one MUFU seed, and the CN sweeps
//    process several partners per iteration for instruction-level parallelism.
#pragma once
#include "d3_molecule.cuh"

namespace d3 {

__device__ __forceinline__ float shx(float v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ double shx(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }

// Transposed reduction over the 16 lanes of a half-warp: on return the lane
// holds the 16-lane total of v[which]; `valid` is false on lanes that hold a
// duplicate.  C <= 16.
template <int C, int OFF, typename S>
struct Red16 {
  static __device__ __forceinline__ S run(const S* v, int hl, int& which, bool& valid) {
    constexpr int NC = (C + 1) / 2;
    const bool bit = (hl & OFF) != 0;
    S nv[NC];
#pragma unroll
    for (int m = 0; m < C / 2; ++m) {
      S keep = bit ? v[2 * m + 1] : v[2 * m];
      S send = bit ? v[2 * m] : v[2 * m + 1];
      nv[m] = keep + shx(send, OFF);
    }
    if (C & 1) nv[NC - 1] = v[C - 1] + shx(v[C - 1], OFF);
    S r = Red16<NC, OFF / 2, S>::run(nv, hl, which, valid);
    if (which < C / 2) {
      which = 2 * which + (bit ? 1 : 0);
    } else {
      which = C - 1;
      if (bit) valid = false;
    }
    return r;
  }
};
template <int C, typename S>
struct Red16<C, 0, S> {
  static __device__ __forceinline__ S run(const S* v, int, int& which, bool& valid) {
    static_assert(C == 1, "Red16 handles at most 16 values");
    which = 0;
    valid = true;
    return v[0];
  }
};

constexpr int kV2MaxWarps = 8;

__host__ __device__ inline int v2_warps(int natoms) {
  int nb = (natoms + 15) / 16;
  int nu = (nb + 1) / 2;
  return nu < 1 ? 1 : (nu > kV2MaxWarps ? kV2MaxWarps : nu);
}

template <typename S>
__host__ __device__ inline size_t molecule_v2_smem_bytes(int natoms) {
  size_t n = (size_t)natoms;
  size_t bytes = n * 24 * sizeof(S);                 // x y z rk q sqrtq cn decn + w[8] + dw[8]
  bytes += n * 5 * v2_warps(natoms) * sizeof(S);     // per-warp accumulators [nw][5][N]
  bytes += n * sizeof(S);                            // g
  bytes += n * 4 * sizeof(int);                      // origZ, sortedZ, perm, group
  bytes += (5 * kMaxZ + 8) * sizeof(int);
  return bytes + 64;
}

template <typename S, bool WANT_G>
__global__ void __launch_bounds__(32 * kV2MaxWarps) molecule_kernel_v2(MolArgs<S> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = A.natoms;
  const int b = blockIdx.x;
  const int t = threadIdx.x;
  const int nt = blockDim.x;
  const int lane = t & 31, wid = t >> 5, nw = nt >> 5;
  const int hl = lane & 15, ph = lane >> 4;

  S* sx = reinterpret_cast<S*>(smem_raw);
  S* sy = sx + N;
  S* sz = sy + N;
  S* srk = sz + N;              // kcn * rcov
  S* sq = srk + N;
  S* ssq = sq + N;
  S* scn = ssq + N;
  S* sdecn = scn + N;
  S* sw = sdecn + N;            // [N][8]
  S* sdw = sw + 8 * (size_t)N;  // [N][8]
  S* acc = sdw + 8 * (size_t)N;  // [nw][5][N]
  S* sg = acc + (size_t)nw * 5 * N;
  int* sorigZ = reinterpret_cast<int*>(sg + N);
  int* sZ = sorigZ + N;
  int* sperm = sZ + N;
  int* sgrp = sperm + N;
  int* cnt = sgrp + N;
  int* zstart = cnt + kMaxZ;
  int* gZ = zstart + kMaxZ;
  int* gstart = gZ + kMaxZ;
  int* zgrp = gstart + kMaxZ + 1;
  int* misc = zgrp + kMaxZ;

  const int64_t* numbers = A.numbers + (size_t)b * N;
  const S* pos = A.positions + (size_t)b * N * 3;

  // ---------------------------------------------------------------- phase 0
  for (int k = t; k < kMaxZ; k += nt) cnt[k] = 0;
  for (int k = t; k < nw * 5 * N; k += nt) acc[k] = S(0);
  __syncthreads();
  for (int k = t; k < N; k += nt) {
    int Z = (int)numbers[k];
    if (Z < 0 || Z >= kMaxZ) Z = 0;
    sorigZ[k] = Z;
    if (Z > 0) {
      atomicAdd(&cnt[Z], 1);
    } else {
      size_t o = (size_t)b * N + k;
      if (A.energy) A.energy[o] = S(0);
      if (WANT_G) { A.grad[3 * o] = S(0); A.grad[3 * o + 1] = S(0); A.grad[3 * o + 2] = S(0); }
    }
  }
  __syncthreads();
  if (t == 0) {
    int ng = 0, run = 0;
    for (int z = 1; z < kMaxZ; ++z) {
      zstart[z] = run;
      if (cnt[z] > 0) { gZ[ng] = z; gstart[ng] = run; zgrp[z] = ng; ++ng; run += cnt[z]; }
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
    sgrp[p] = zgrp[Z];
    sx[p] = pos[3 * k]; sy[p] = pos[3 * k + 1]; sz[p] = pos[3 * k + 2];
    S rc = A.rcov_per_element ? A.rcov[Z] : A.rcov[(size_t)b * N + k];
    S q = A.r4r2_per_element ? A.r4r2[Z] : A.r4r2[(size_t)b * N + k];
    srk[p] = A.kcn * rc;
    sq[p] = q;
    ssq[p] = sqr(q);
    sg[p] = (WANT_G && A.g) ? A.g[(size_t)b * N + k] : S(1);
  }
  __syncthreads();

  const S eps = eps_of<S>();
  const int nb = (n + 15) >> 4;
  const int nunits = (nb + 1) >> 1;
  S* accw = acc + (size_t)wid * 5 * N;

  // ---------------------------------------------------------------- phase 1
  // CN_i = sum_j [r_ij <= 25] / (1 + exp(-kcn (rc_ij / r_ij - 1))), pairs once
  constexpr int CNB = 4;
  for (int u = wid; u < nunits; u += nw) {
    for (int side = 0; side < 2; ++side) {
      const int rb = side == 0 ? u : nb - 1 - u;
      if (side == 1 && rb == u) break;
      const int i0 = rb << 4, i = i0 + hl;
      const bool irow = i < n;
      const int ic = irow ? i : n - 1;
      const S xi = sx[ic], yi = sy[ic], zi = sz[ic], rki = srk[ic];
      S cni = S(0);
      for (int jb0 = i0 + 1; jb0 < n; jb0 += 2 * CNB) {  // warp-uniform trip count
        const int jb = jb0 + ph;
        S f[CNB];
#pragma unroll
        for (int s = 0; s < CNB; ++s) {
          const int j = jb + 2 * s;
          const int jc = j < n ? j : n - 1;
          S dx = xi - sx[jc], dy = yi - sy[jc], dz = zi - sz[jc];
          S r2 = dx * dx + dy * dy + dz * dz;
          const bool ok = irow && j < n && j > i && r2 <= A.cn_cutoff2;
          r2 = fmx(r2, eps);
          S rinv = frsq(r2);
          S x = fmx(A.kcn - (rki + srk[jc]) * rinv, S(-700));
          S v = frcp(S(1) + fex(x));
          f[s] = ok ? v : S(0);
          cni += f[s];
        }
        int which; bool valid;
        S tot = Red16<CNB, 8, S>::run(f, hl, which, valid);
        const int j = jb + 2 * which;
        if (valid && j < n) accw[j] += tot;
      }
      cni += shx(cni, 16);
      __syncwarp();
      if (ph == 0 && irow) accw[i] += cni;
      __syncwarp();
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase 2
  for (int i = t; i < n; i += nt) {
    S cn = S(0);
    for (int w = 0; w < nw; ++w) { cn += acc[(size_t)w * 5 * N + i]; acc[(size_t)w * 5 * N + i] = S(0); }
    scn[i] = cn;
    S wi[kNRef], dwi[kNRef];
    reference_weights<S>(A.refcn + sZ[i] * kNRef, cn, wi, dwi);
#pragma unroll
    for (int a = 0; a < kNRef; ++a) { sw[8 * i + a] = wi[a]; sdw[8 * i + a] = dwi[a]; }
    sw[8 * i + 7] = S(0); sdw[8 * i + 7] = S(0);
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase 3
  for (int u = wid; u < nunits; u += nw) {
    for (int side = 0; side < 2; ++side) {
      const int rb = side == 0 ? u : nb - 1 - u;
      if (side == 1 && rb == u) break;
      const int i0 = rb << 4, i = i0 + hl;
      const bool irow = i < n;
      const int ic = irow ? i : n - 1;
      const S xi = sx[ic], yi = sy[ic], zi = sz[ic], qi = sq[ic];
      const int Zi = sZ[ic];
      const S gi = sg[ic];
      const S sa_i = A.a1 * sqr(S(3) * qi);
      const S q3_i = S(3) * qi;
      S ei = S(0), gxi = S(0), gyi = S(0), gzi = S(0), di = S(0);
      for (int grp = sgrp[i0]; grp < ng; ++grp) {
        const int jlo = max(gstart[grp], i0 + 1), jhi = gstart[grp + 1];
        if (jlo >= jhi) continue;
        // U[b] = sum_a w_ia K[Zi,Zg,a,b],  Ud[b] = sum_a dw_ia K[Zi,Zg,a,b]:
        // half 0 contracts the weights, half 1 their CN derivatives, then swap
        const S* __restrict__ K = A.refc6 + ((size_t)Zi * kMaxZ + gZ[grp]) * (kNRef * kNRef);
        S U[kNRef], Ud[kNRef];
        {
          const S* src = (WANT_G && ph == 1) ? sdw + 8 * ic : sw + 8 * ic;
          S mine[kNRef];
#pragma unroll
          for (int bb = 0; bb < kNRef; ++bb) mine[bb] = S(0);
#pragma unroll
          for (int a = 0; a < kNRef; ++a) {
            const S wa = src[a];
#pragma unroll
            for (int bb = 0; bb < kNRef; ++bb) mine[bb] += wa * __ldg(K + a * kNRef + bb);
          }
#pragma unroll
          for (int bb = 0; bb < kNRef; ++bb) {
            if (WANT_G) {
              S other = shx(mine[bb], 16);
              U[bb] = ph == 0 ? mine[bb] : other;
              Ud[bb] = ph == 0 ? other : mine[bb];
            } else {
              U[bb] = mine[bb];
            }
          }
        }
        for (int j = jlo + ph; j < jhi + ph; j += 2) {
          // (the loop bound keeps both halves in step; the last j of half 1 may be jhi)
          const bool jin = j < jhi;
          const int jc = jin ? j : jhi - 1;
          S dx = xi - sx[jc], dy = yi - sy[jc], dz = zi - sz[jc];
          S r2 = dx * dx + dy * dy + dz * dz;
          const bool ok = irow && jin && jc > i && r2 <= A.cutoff2;
          r2 = fmx(r2, eps);
          const S* wj = sw + 8 * jc;
          S c6 = U[0] * wj[0];
#pragma unroll
          for (int a = 1; a < kNRef; ++a) c6 += U[a] * wj[a];
          S qq = q3_i * sq[jc];
          S r0 = sa_i * ssq[jc] + A.a2;
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
          S P = A.s6 * t6 + s8q * t8;
          S vals[5];
          vals[0] = ok ? c6 * P : S(0);
          ei += vals[0];
          if (WANT_G) {
            const S* dwj = sdw + 8 * jc;
            S dci = Ud[0] * wj[0], dcj = U[0] * dwj[0];
#pragma unroll
            for (int a = 1; a < kNRef; ++a) { dci += Ud[a] * wj[a]; dcj += U[a] * dwj[a]; }
            S gam = ok ? S(0.5) * (gi + sg[jc]) : S(0);
            S gP = gam * P;
            di -= gP * dci;
            S t62 = t6 * t6, t82 = t8 * t8;
            S dP = S(3) * (A.s6 * (r4 * t62)) + S(4) * (s8q * (r6 * t82));
            S f = (S(2) * gam) * (c6 * dP);
            S fx = f * dx, fy = f * dy, fz = f * dz;
            gxi += fx; gyi += fy; gzi += fz;
            vals[1] = -fx; vals[2] = -fy; vals[3] = -fz;
            vals[4] = -(gP * dcj);
          } else {
            vals[1] = vals[2] = vals[3] = vals[4] = S(0);
          }
          int which; bool valid;
          if (WANT_G) {
            S tot = Red16<5, 8, S>::run(vals, hl, which, valid);
            if (valid && jin) accw[(size_t)which * N + jc] += tot;
          } else {
            S tot = Red16<1, 8, S>::run(vals, hl, which, valid);
            if (valid && jin) accw[jc] += tot;
          }
        }
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
  }
  __syncthreads();
  for (int i = t; i < n; i += nt) {
    S e = S(0), d = S(0);
    for (int w = 0; w < nw; ++w) { e += acc[(size_t)w * 5 * N + i]; d += acc[((size_t)w * 5 + 4) * N + i]; }
    if (A.energy) A.energy[(size_t)b * N + sperm[i]] = S(-0.5) * e;
    sdecn[i] = d;
  }
  if (!WANT_G) return;
  __syncthreads();

  // ---------------------------------------------------------------- phase 4
  // CN chain: (dL/dCN_i + dL/dCN_j) df_ij/dx, df/dx_i = -kcn rc r^-3 f(1-f) (x_i - x_j)
  constexpr int GB = 2;
  for (int u = wid; u < nunits; u += nw) {
    for (int side = 0; side < 2; ++side) {
      const int rb = side == 0 ? u : nb - 1 - u;
      if (side == 1 && rb == u) break;
      const int i0 = rb << 4, i = i0 + hl;
      const bool irow = i < n;
      const int ic = irow ? i : n - 1;
      const S xi = sx[ic], yi = sy[ic], zi = sz[ic], rki = srk[ic], dci = sdecn[ic];
      S gxi = S(0), gyi = S(0), gzi = S(0);
      for (int jb0 = i0 + 1; jb0 < n; jb0 += 2 * GB) {  // warp-uniform trip count
        const int jb = jb0 + ph;
        S vals[3 * GB];
#pragma unroll
        for (int s = 0; s < GB; ++s) {
          const int j = jb + 2 * s;
          const int jc = j < n ? j : n - 1;
          S dx = xi - sx[jc], dy = yi - sy[jc], dz = zi - sz[jc];
          S r2 = dx * dx + dy * dy + dz * dz;
          const bool ok = irow && j < n && j > i && r2 <= A.cn_cutoff2 && r2 >= eps;
          r2 = fmx(r2, eps);
          S rinv = frsq(r2);
          S rk = rki + srk[jc];
          S e = fex(fmx(A.kcn - rk * rinv, S(-700)));
          S f = frcp(S(1) + e);
          S df = -(rk * (rinv * rinv * rinv)) * (f * (e * f));
          S c = ok ? (dci + sdecn[jc]) * df : S(0);
          S cx = c * dx, cy = c * dy, cz = c * dz;
          gxi += cx; gyi += cy; gzi += cz;
          vals[3 * s] = -cx; vals[3 * s + 1] = -cy; vals[3 * s + 2] = -cz;
        }
        int which; bool valid;
        S tot = Red16<3 * GB, 8, S>::run(vals, hl, which, valid);
        const int j = jb + 2 * (which / 3);
        if (valid && j < n) accw[(size_t)(1 + which % 3) * N + j] += tot;
      }
      gxi += shx(gxi, 16); gyi += shx(gyi, 16); gzi += shx(gzi, 16);
      __syncwarp();
      if (ph == 0 && irow) { accw[N + i] += gxi; accw[2 * N + i] += gyi; accw[3 * N + i] += gzi; }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = t; i < n; i += nt) {
    S gx = S(0), gy = S(0), gz = S(0);
    for (int w = 0; w < nw; ++w) {
      const S* a = acc + (size_t)w * 5 * N;
      gx += a[N + i]; gy += a[2 * N + i]; gz += a[3 * N + i];
    }
    size_t o = (size_t)b * N + sperm[i];
    A.grad[3 * o] = gx; A.grad[3 * o + 1] = gy; A.grad[3 * o + 2] = gz;
  }
}

}  // namespace d3
