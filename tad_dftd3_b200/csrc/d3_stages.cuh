// Stage-level kernels: the reference's "operator API" (README.rst:238-261) keeps
// the intermediates -- cn [B,N], weights [B,N,7], dense c6 [B,N,N] -- as tensors
// and lets the caller feed its own dense c6 / rvdw to `dispersion`.  These
// kernels reproduce each stage on its own, forward and vector-Jacobian product (the `*_bwd`
// kernels: what torch autograd derives for the reference's stage functions, README.rst:238-261
// under `autograd.grad`); the fused kernels never materialise the dense tensors.
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
                                S cutoff2, S kcn, S* cn, int counting = kCountExp) {
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
      if (counting == kCountExp) {
        S rinv = rsq(r2);
        acc += rcp(S(1) + ex(-kcn * ((ri + rcov[o + j]) * rinv - S(1))));
      } else {
        S f, df;
        counting_pair<S, S>(counting, r2, ri + rcov[o + j], kcn, f, df);
        acc += f;
      }
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
                                   const S* r4r2, S s6, S s8, S a1, S a2, S cutoff2, S* energy,
                                   int damping = kDampRational, S alp = S(14)) {
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
      S c = c6[(o + i) * natoms + j];
      if (damping == kDampRational) {
        S r0 = a1 * sqr(qq) + a2;
        S r02 = r0 * r0, r06 = r02 * r02 * r02, r08 = r06 * r02;
        S r4 = r2 * r2, r6 = r4 * r2, r8 = r4 * r4;
        e6 += c * rcp(r6 + r06);
        e8 += (c * qq) * rcp(r8 + r08);
      } else {
        S t6, t8, dt6, dt8, t6p1, t6p2, t8p1, t8p2;
        damping_pair<S, S>(damping, r2, sqr(qq), a1, a2, alp, t6, t8, dt6, dt8, t6p1, t6p2, t8p1, t8p2);
        e6 += c * t6;
        e8 += (c * qq) * t8;
      }
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


// ---------------------------------------------------------------------------------------------
// Vector-Jacobian products of the stage functions (first order; cotangents in, gradients out).
// ---------------------------------------------------------------------------------------------

// cn_d3: gx_i = sum_j (gcn_i + gcn_j) f'(r_ij) (x_i - x_j) / r_ij       (thread per atom, all partners)
template <typename S>
__global__ void stage_cn_bwd_kernel(int natoms, const int64_t* numbers, const S* pos, const S* rcov,
                                    S cutoff2, S kcn, const S* gcn, S* gx, int counting = kCountExp) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natoms) return;
  const size_t o = (size_t)b * natoms;
  S ax = S(0), ay = S(0), az = S(0);
  if (numbers[o + i] != 0) {
    const S xi = pos[3 * (o + i)], yi = pos[3 * (o + i) + 1], zi = pos[3 * (o + i) + 2], ri = rcov[o + i];
    const S gi = gcn[o + i];
    const S eps = eps_of<S>();
    for (int j = 0; j < natoms; ++j) {
      if (j == i || numbers[o + j] == 0) continue;
      S dx = xi - pos[3 * (o + j)], dy = yi - pos[3 * (o + j) + 1], dz = zi - pos[3 * (o + j) + 2];
      S r2 = dx * dx + dy * dy + dz * dz;
      if (r2 > cutoff2 || r2 < eps) continue;          // clamped distance: no position dependence
      S f, df;
      counting_pair<S, S>(counting, r2, ri + rcov[o + j], kcn, f, df);
      const S c = (gi + gcn[o + j]) * df;
      ax += c * dx; ay += c * dy; az += c * dz;
    }
  }
  gx[3 * (o + i)] = ax; gx[3 * (o + i) + 1] = ay; gx[3 * (o + i) + 2] = az;
}

// weight_references: gcn_i = sum_a gw_ia dw_ia / dCN_i
template <typename S>
__global__ void stage_weights_bwd_kernel(size_t n, const int64_t* numbers, const S* cn, const S* refcn,
                                         const S* gw, S* gcn) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int Z = (int)numbers[i];
  if (Z < 0 || Z >= kMaxZ) Z = 0;
  S wi[kNRef], dwi[kNRef];
  reference_weights<S>(refcn + Z * kNRef, cn[i], wi, dwi);
  S acc = S(0);
#pragma unroll
  for (int a = 0; a < kNRef; ++a) acc += gw[i * kNRef + a] * dwi[a];
  gcn[i] = Z > 0 ? acc : S(0);
}

// dispersion2 with a dense c6 (E_i = -1/2 sum_j c6_ij P_ij, entries of c6 independent):
//   gc6_ij = -1/2 ge_i P_ij,   gx_i = sum_j (ge_i c6_ij + ge_j c6_ji) (-dP/dr^2)_ij (x_i - x_j),
//   gpar_i[4] = ge_i dE_i / d(s6, s8, a1, a2)                                (summed over atoms by the caller)
template <typename S>
__global__ void stage_disp2_bwd_kernel(int natoms, const int64_t* numbers, const S* pos, const S* c6,
                                       const S* r4r2, S s6, S s8, S a1, S a2, S cutoff2, const S* ge,
                                       S* gx, S* gc6, S* gpar, int damping = kDampRational, S alp = S(14)) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natoms) return;
  const size_t o = (size_t)b * natoms;
  S ax = S(0), ay = S(0), az = S(0), p6 = S(0), p8 = S(0), p1 = S(0), p2 = S(0);
  const bool real_i = numbers[o + i] != 0;
  const S xi = pos[3 * (o + i)], yi = pos[3 * (o + i) + 1], zi = pos[3 * (o + i) + 2];
  const S qi = r4r2[o + i], gi = ge[o + i];
  const S eps = eps_of<S>();
  for (int j = 0; j < natoms; ++j) {
    S gcij = S(0);
    if (real_i && j != i && numbers[o + j] != 0) {
      S dx = xi - pos[3 * (o + j)], dy = yi - pos[3 * (o + j) + 1], dz = zi - pos[3 * (o + j) + 2];
      S r2 = dx * dx + dy * dy + dz * dz;
      if (r2 <= cutoff2) {
        const bool clamped = r2 < eps;
        if (clamped) r2 = eps;
        const S qq = S(3) * qi * r4r2[o + j];
        S t6, t8, dt6, dt8, t6p1, t6p2, t8p1, t8p2;
        damping_pair<S, S>(damping, r2, sqr(qq), a1, a2, alp, t6, t8, dt6, dt8, t6p1, t6p2, t8p1, t8p2);
        const S cij = c6[(o + i) * natoms + j], cji = c6[(o + j) * natoms + i];
        gcij = S(-0.5) * gi * (s6 * t6 + s8 * qq * t8);
        if (!clamped) {
          const S f = (gi * cij + ge[o + j] * cji) * (s6 * dt6 + s8 * qq * dt8);
          ax += f * dx; ay += f * dy; az += f * dz;
        }
        const S h = S(-0.5) * gi * cij;
        p6 += h * t6;
        p8 += h * (qq * t8);
        p1 += h * (s6 * t6p1 + s8 * qq * t8p1);
        p2 += h * (s6 * t6p2 + s8 * qq * t8p2);
      }
    }
    gc6[(o + i) * natoms + j] = gcij;
  }
  gx[3 * (o + i)] = ax; gx[3 * (o + i) + 1] = ay; gx[3 * (o + i) + 2] = az;
  gpar[4 * (o + i)] = p6; gpar[4 * (o + i) + 1] = p8; gpar[4 * (o + i) + 2] = p1; gpar[4 * (o + i) + 3] = p2;
}

// e / c9 of one triple and its derivatives with respect to a = r_ij^2, b = r_ik^2, c = r_jk^2
// (damping/atm.py:113-146; symmetric in a, b, c)
template <typename S>
__device__ __forceinline__ S atm_core3(S a, S b, S c, S r0, S pexp, S& da, S& db, S& dc) {
  const S p2 = a * b * c;
  const S p1i = rsq(p2);                          // 1 / (r_ij r_ik r_jk)
  const S p3i = p1i * p1i * p1i, p5i = p3i * p1i * p1i;
  const S u1 = a + c - b, u2 = a - c + b, u3 = -a + c + b;
  const S s = u1 * u2 * u3;
  const S ang = S(0.375) * s * p5i + p3i;
  const S t = pw(r0 * p1i, pexp);
  const S fd = rcp(S(1) + S(6) * t);
  const S rad = S(1.875) * s * p5i + S(3) * p3i;
  const S dfd = S(3) * pexp * t * fd * fd;
  da = (S(0.375) * (u2 * u3 + u1 * u3 - u1 * u2) * p5i - rad * (S(0.5) * rcp(a))) * fd + ang * dfd * rcp(a);
  db = (S(0.375) * (-(u2 * u3) + u1 * u3 + u1 * u2) * p5i - rad * (S(0.5) * rcp(b))) * fd + ang * dfd * rcp(b);
  dc = (S(0.375) * (u2 * u3 - u1 * u3 + u1 * u2) * p5i - rad * (S(0.5) * rcp(c))) * fd + ang * dfd * rcp(c);
  return ang * fd;
}

// dispersion_atm with dense c6 / rvdw: thread per ordered pair (i, j) with r_ij <= cutoff, loop over k.
// The ordered term (i; j, k) uses the entries c6[i,j], c6[i,k], c6[j,k] exactly as the reference indexes them;
// gx, gc6 must be zero on entry (fp64 atomics).
template <typename S>
__global__ void stage_atm_bwd_kernel(int natoms, const int64_t* numbers, const S* pos, const S* c6,
                                     const S* rvdw, S cutoff2, S s9, S rs9, S pexp, const S* ge, S* gx, S* gc6) {
  const int b = blockIdx.z, i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= natoms || j == i) return;
  const size_t o = (size_t)b * natoms;
  if (numbers[o + i] == 0 || numbers[o + j] == 0) return;
  const S eps = eps_of<S>();
  const S xi = pos[3 * (o + i)], yi = pos[3 * (o + i) + 1], zi = pos[3 * (o + i) + 2];
  const S xj = pos[3 * (o + j)], yj = pos[3 * (o + j) + 1], zj = pos[3 * (o + j) + 2];
  const S dxa = xi - xj, dya = yi - yj, dza = zi - zj;
  S a = dxa * dxa + dya * dya + dza * dza;
  if (a < eps) a = eps;
  if (a > cutoff2) return;
  const S w = ge[o + i] / S(6);
  const S rs93 = rs9 * rs9 * rs9;
  const S* c6i = c6 + (o + i) * natoms;
  const S* c6j = c6 + (o + j) * natoms;
  const S* rvi = rvdw + (o + i) * natoms;
  const S* rvj = rvdw + (o + j) * natoms;
  const S cij = c6i[j];
  S gix = S(0), giy = S(0), giz = S(0), gjx = S(0), gjy = S(0), gjz = S(0), gcij = S(0);
  for (int k = 0; k < natoms; ++k) {
    if (k == i || k == j || numbers[o + k] == 0) continue;
    const S xk = pos[3 * (o + k)], yk = pos[3 * (o + k) + 1], zk = pos[3 * (o + k) + 2];
    const S dxc = xj - xk, dyc = yj - yk, dzc = zj - zk;
    S c = dxc * dxc + dyc * dyc + dzc * dzc;
    if (c < eps) c = eps;
    if (c > cutoff2) continue;
    const S dxb = xi - xk, dyb = yi - yk, dzb = zi - zk;
    S bq = dxb * dxb + dyb * dyb + dzb * dzb;
    if (bq < eps) bq = eps;
    const S cik = c6i[k], cjk = c6j[k];
    const S prod = ab(cij * cik * cjk);
    const bool clamped = prod < eps;
    const S c9 = s9 * sqr(clamped ? eps : prod);
    const S r0 = rs93 * (rvi[j] * rvi[k] * rvj[k]);
    S da, db, dc;
    const S core = atm_core3<S>(a, bq, c, r0, pexp, da, db, dc);
    const S fa = S(2) * w * c9 * da, fb = S(2) * w * c9 * db, fc = S(2) * w * c9 * dc;
    gix += fa * dxa + fb * dxb; giy += fa * dya + fb * dyb; giz += fa * dza + fb * dzb;
    gjx += -fa * dxa + fc * dxc; gjy += -fa * dya + fc * dyc; gjz += -fa * dza + fc * dzc;
    red_add(gx + 3 * (o + k), -fb * dxb - fc * dxc);
    red_add(gx + 3 * (o + k) + 1, -fb * dyb - fc * dyc);
    red_add(gx + 3 * (o + k) + 2, -fb * dzb - fc * dzc);
    if (!clamped) {
      const S h = S(0.5) * w * c9 * core;
      gcij += h * rcp(cij);
      red_add(gc6 + (o + i) * natoms + k, h * rcp(cik));
      red_add(gc6 + (o + j) * natoms + k, h * rcp(cjk));
    }
  }
  red_add(gx + 3 * (o + i), gix); red_add(gx + 3 * (o + i) + 1, giy); red_add(gx + 3 * (o + i) + 2, giz);
  red_add(gx + 3 * (o + j), gjx); red_add(gx + 3 * (o + j) + 1, gjy); red_add(gx + 3 * (o + j) + 2, gjz);
  red_add(gc6 + (o + i) * natoms + j, gcij);
}

}  // namespace d3
