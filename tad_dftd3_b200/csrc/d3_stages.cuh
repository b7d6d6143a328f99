// Stage-level kernels: the reference's "operator API" (README.rst:238-261) keeps
// the intermediates -- cn [B,N], weights [B,N,7], dense c6 [B,N,N] -- as tensors
// and lets the caller feed its own dense c6 / rvdw to `dispersion`.  These
// kernels reproduce each stage on its own (forward only); the fused kernels
// never materialise the dense tensors.
//   stage_cn       tad-mctc cn_d3 + exp_count          (call site disp.py:150-152)
//   stage_weights  model/weights.py:104-176
//   stage_c6       model/c6.py:186-217
//   stage_disp2    disp.py:276-302 + damping/rational.py:68   (dense c6 input)
//   stage_atm      damping/atm.py:86-155                      (dense c6, rvdw input)
#pragma once
#include "d3_molecule.cuh"

namespace d3 {

template <typename S>
__global__ void stage_cn_kernel(int natoms, const int64_t* numbers, const S* pos, const S* rcov,
                                S cutoff2, S kcn, S* cn) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natoms) return;
  const size_t o = (size_t)b * natoms;
  S acc = S(0);
  if (numbers[o + i] != 0) {
    const S xi = pos[3 * (o + i)], yi = pos[3 * (o + i) + 1], zi = pos[3 * (o + i) + 2], ri = rcov[o + i];
    const S eps = eps_of<S>();
    for (int j = 0; j < natoms; ++j) {
      if (j == i || numbers[o + j] == 0) continue;
      S dx = xi - pos[3 * (o + j)], dy = yi - pos[3 * (o + j) + 1], dz = zi - pos[3 * (o + j) + 2];
      S r2 = dx * dx + dy * dy + dz * dz;
      if (r2 > cutoff2) continue;
      if (r2 < eps) r2 = eps;
      S rinv = rsq(r2);
      acc += rcp(S(1) + ex(-kcn * ((ri + rcov[o + j]) * rinv - S(1))));
    }
  }
  cn[o + i] = acc;
}

template <typename S>
__global__ void stage_weights_kernel(size_t n, const int64_t* numbers, const S* cn, const S* refcn, S* w) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int Z = (int)numbers[i];
  if (Z < 0 || Z >= kMaxZ) Z = 0;
  S wi[kNRef], dwi[kNRef];
  reference_weights<S>(refcn + Z * kNRef, cn[i], wi, dwi);   // Z = 0: every reference absent -> zeros
#pragma unroll
  for (int a = 0; a < kNRef; ++a) w[i * kNRef + a] = wi[a];
}

template <typename S>
__global__ void stage_c6_kernel(int natoms, const int64_t* numbers, const S* w, const S* refc6, S* c6) {
  const int b = blockIdx.z, i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= natoms) return;
  const size_t o = (size_t)b * natoms;
  int Zi = (int)numbers[o + i], Zj = (int)numbers[o + j];
  if (Zi < 0 || Zi >= kMaxZ) Zi = 0;
  if (Zj < 0 || Zj >= kMaxZ) Zj = 0;
  const S* K = refc6 + ((size_t)Zi * kMaxZ + Zj) * (kNRef * kNRef);
  const S* wi = w + (o + i) * kNRef;
  const S* wj = w + (o + j) * kNRef;
  S acc = S(0);
#pragma unroll
  for (int a = 0; a < kNRef; ++a) {
    S t = S(0);
#pragma unroll
    for (int bb = 0; bb < kNRef; ++bb) t += K[a * kNRef + bb] * wj[bb];
    acc += wi[a] * t;
  }
  c6[(o + i) * natoms + j] = acc;
}

// Bilinear form behind atomic_c6 and its derivatives: out_ij = sum_ab u_ia K[Zi,Zj,a,b] w_jb
// (atomic_c6 is u = w; the cotangent of `stage_c6_contract_kernel` is u = v).
template <typename S>
__global__ void stage_c6_bilinear_kernel(int natoms, const int64_t* numbers, const S* u, const S* w,
                                         const S* refc6, S* out) {
  const int b = blockIdx.z, i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= natoms) return;
  const size_t o = (size_t)b * natoms;
  int Zi = (int)numbers[o + i], Zj = (int)numbers[o + j];
  if (Zi < 0 || Zi >= kMaxZ) Zi = 0;
  if (Zj < 0 || Zj >= kMaxZ) Zj = 0;
  const S* K = refc6 + ((size_t)Zi * kMaxZ + Zj) * (kNRef * kNRef);
  const S* ui = u + (o + i) * kNRef;
  const S* wj = w + (o + j) * kNRef;
  S acc = S(0);
#pragma unroll
  for (int a = 0; a < kNRef; ++a) {
    S t = S(0);
#pragma unroll
    for (int bb = 0; bb < kNRef; ++bb) t += K[a * kNRef + bb] * wj[bb];
    acc += ui[a] * t;
  }
  out[(o + i) * natoms + j] = acc;
}

// Contraction behind the backward of atomic_c6 (reference model/c6.py:317-335, the
// einsums "...ijab,...jb->...ija" and "...ij,...ija->...ia" in one pass):
//   out_ia = sum_j G_ij sum_b K[Zi,Zj,a,b] w_jb          one thread per (structure, i)
template <typename S>
__global__ void stage_c6_contract_kernel(int natoms, const int64_t* numbers, const S* G, const S* w,
                                         const S* refc6, S* out) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natoms) return;
  const size_t o = (size_t)b * natoms;
  int Zi = (int)numbers[o + i];
  if (Zi < 0 || Zi >= kMaxZ) Zi = 0;
  S acc[kNRef];
#pragma unroll
  for (int a = 0; a < kNRef; ++a) acc[a] = S(0);
  for (int j = 0; j < natoms; ++j) {
    int Zj = (int)numbers[o + j];
    if (Zj <= 0 || Zj >= kMaxZ) continue;
    const S g = G[(o + i) * natoms + j];
    const S* K = refc6 + ((size_t)Zi * kMaxZ + Zj) * (kNRef * kNRef);
    const S* wj = w + (o + j) * kNRef;
    S wv[kNRef];
#pragma unroll
    for (int bb = 0; bb < kNRef; ++bb) wv[bb] = wj[bb];
#pragma unroll
    for (int a = 0; a < kNRef; ++a) {
      S t = S(0);
#pragma unroll
      for (int bb = 0; bb < kNRef; ++bb) t += K[a * kNRef + bb] * wv[bb];
      acc[a] += g * t;
    }
  }
#pragma unroll
  for (int a = 0; a < kNRef; ++a) out[(o + i) * kNRef + a] = Zi > 0 ? acc[a] : S(0);
}

template <typename S>
__global__ void stage_disp2_kernel(int natoms, const int64_t* numbers, const S* pos, const S* c6,
                                   const S* r4r2, S s6, S s8, S a1, S a2, S cutoff2, S* energy) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natoms) return;
  const size_t o = (size_t)b * natoms;
  S e6 = S(0), e8 = S(0);
  if (numbers[o + i] != 0) {
    const S xi = pos[3 * (o + i)], yi = pos[3 * (o + i) + 1], zi = pos[3 * (o + i) + 2], qi = r4r2[o + i];
    const S eps = eps_of<S>();
    for (int j = 0; j < natoms; ++j) {
      if (j == i || numbers[o + j] == 0) continue;
      S dx = xi - pos[3 * (o + j)], dy = yi - pos[3 * (o + j) + 1], dz = zi - pos[3 * (o + j) + 2];
      S r2 = dx * dx + dy * dy + dz * dz;
      if (r2 > cutoff2) continue;
      if (r2 < eps) r2 = eps;
      S qq = S(3) * qi * r4r2[o + j];
      S r0 = a1 * sqr(qq) + a2;
      S r02 = r0 * r0, r06 = r02 * r02 * r02, r08 = r06 * r02;
      S r4 = r2 * r2, r6 = r4 * r2, r8 = r4 * r4;
      S c = c6[(o + i) * natoms + j];
      e6 += c * rcp(r6 + r06);
      e8 += (c * qq) * rcp(r8 + r08);
    }
  }
  energy[o + i] = s6 * (S(-0.5) * e6) + s8 * (S(-0.5) * e8);
}

// dense ordered ATM sum, thread per i (stage use only: small structures)
template <typename S>
__global__ void stage_atm_kernel(int natoms, const int64_t* numbers, const S* pos, const S* c6,
                                 const S* rvdw, S cutoff2, S s9, S rs9, S pexp, S* energy) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natoms) return;
  const size_t o = (size_t)b * natoms;
  S acc = S(0);
  if (numbers[o + i] != 0) {
    const S eps = eps_of<S>();
    const S xi = pos[3 * (o + i)], yi = pos[3 * (o + i) + 1], zi = pos[3 * (o + i) + 2];
    const S rs93 = rs9 * rs9 * rs9;
    const S* c6i = c6 + (o + i) * natoms;
    const S* rvi = rvdw + (o + i) * natoms;
    for (int j = 0; j < natoms; ++j) {
      if (j == i || numbers[o + j] == 0) continue;
      const S xj = pos[3 * (o + j)], yj = pos[3 * (o + j) + 1], zj = pos[3 * (o + j) + 2];
      S a = (xi - xj) * (xi - xj) + (yi - yj) * (yi - yj) + (zi - zj) * (zi - zj);
      if (a < eps) a = eps;
      if (a > cutoff2) continue;                       // [r_ij <= c]
      const S* c6j = c6 + (o + j) * natoms;
      const S* rvj = rvdw + (o + j) * natoms;
      for (int k = 0; k < natoms; ++k) {
        if (k == i || k == j || numbers[o + k] == 0) continue;
        const S xk = pos[3 * (o + k)], yk = pos[3 * (o + k) + 1], zk = pos[3 * (o + k) + 2];
        S c = (xj - xk) * (xj - xk) + (yj - yk) * (yj - yk) + (zj - zk) * (zj - zk);
        if (c < eps) c = eps;
        if (c > cutoff2) continue;                     // [r_jk <= c]; r_ik is not tested
        S bq = (xi - xk) * (xi - xk) + (yi - yk) * (yi - yk) + (zi - zk) * (zi - zk);
        if (bq < eps) bq = eps;
        S prod = ab(c6i[j] * c6i[k] * c6j[k]);
        S c9 = s9 * sqr(prod < eps ? eps : prod);
        S r0 = rs93 * (rvi[j] * rvi[k] * rvj[k]);
        S e, de;
        atm_term<S, S>(a, bq, c, rcp(a), c9, r0, pexp, e, de);
        acc += e;
      }
    }
  }
  energy[o + i] = acc / S(6);
}

}  // namespace d3
