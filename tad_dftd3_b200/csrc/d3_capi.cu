// C ABI of libd3b200.so (declared in include/d3b200.h).
// Argument checking, launch configuration and template dispatch only; all
// arithmetic lives in the kernel headers.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/d3b200.h"
#include "d3_molecule.cuh"
#include "d3_molecule_v2.cuh"
#include "d3_system.cuh"
#include "d3_atm.cuh"
#include "d3_atm3.cuh"
#include "d3_stages.cuh"
#include "d3_system_sym.cuh"
// 8-row-group form of the pair sweep: measured SLOWER than the 32-row form at the 50-Bohr cutoff (3.12 against
// 2.61 ms on the 50k-atom cluster: the per-step butterfly, four distinct T rows per step and four candidate lists per
// tile cost more than the lane efficiency gains, 59 -> 71 %); kept behind this switch, off by default.
#ifndef D3_SYM_PAIR_Q8
#define D3_SYM_PAIR_Q8 0
#endif

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* a = "", long b = 0, long c = 0) {
  snprintf(g_err, sizeof(g_err), fmt, a, b, c);
  return code;
}

int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "CUDA error in %s: %s", where, cudaGetErrorString(e));
  return D3B200_ECUDA;
}

constexpr size_t kSmemLimit = 227 * 1024;

template <typename S>
int max_atoms() {
  int n = 4096;
  while (n > 32 && d3::molecule_smem_bytes<S>(n) > kSmemLimit) n -= 32;
  return n;
}

template <typename B>
struct ModelOf;
template <> struct ModelOf<double> { using type = d3b200_model_f64; };
template <> struct ModelOf<float> { using type = d3b200_model_f32; };

template <typename S, int FLAGS>
int launch(const d3::MolArgs<S>& args, cudaStream_t stream) {
  const int N = args.natoms;
  const int threads = N >= d3::kMolThreads ? d3::kMolThreads : ((N + 31) / 32) * 32;
  const size_t smem = d3::molecule_smem_bytes<S>(N);
  auto kern = d3::molecule_kernel<S, FLAGS>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
  }
  kern<<<args.nbatch, threads, smem, stream>>>(args);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "molecule_kernel launch");
  return D3B200_OK;
}

// second-generation kernel (pairs evaluated once): energies / first derivatives,
// two-body + CN, float or double
template <typename B, bool WANT_G>
int launch_v2(const d3::MolArgs<B>& args, cudaStream_t stream) {
  const int N = args.natoms;
  const int threads = 32 * d3::v2_warps(N);
  const size_t smem = d3::molecule_v2_smem_bytes<B>(N);
  auto kern = args.r4r2_per_element ? d3::molecule_kernel_v2<B, WANT_G, true> : d3::molecule_kernel_v2<B, WANT_G, false>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
  }
  kern<<<args.nbatch, threads, smem, stream>>>(args);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "molecule_kernel_v2 launch");
  return D3B200_OK;
}

template <typename B>
bool v2_fits(int natoms) { return d3::molecule_v2_smem_bytes<B>(natoms) <= kSmemLimit; }

template <typename S, typename B>
int fill_common(d3::MolArgs<S>& a, int nbatch, int natoms, const int64_t* numbers,
                const B* positions, const typename ModelOf<B>::type* m,
                const d3b200_params* par, const d3b200_params* dpar) {
  if (nbatch < 0 || natoms < 0) return fail(D3B200_EINVAL, "negative size%s");
  if (!numbers || !positions || !m || !par) return fail(D3B200_EINVAL, "null argument%s");
  if (!m->rcov || !m->r4r2 || !m->refcn || !m->refc6) return fail(D3B200_EINVAL, "null model table%s");
  if (natoms > max_atoms<S>())
    return fail(D3B200_ESIZE, "natoms = %s%ld exceeds the fused-kernel limit %ld", "", natoms, max_atoms<S>());
  if (par->s9 != 0.0 && (m->rvdw_mode == 0 || !m->rvdw))
    return fail(D3B200_EINVAL, "s9 != 0 needs rvdw%s");
  memset(&a, 0, sizeof(a));
  a.nbatch = nbatch;
  a.natoms = natoms;
  a.numbers = numbers;
  a.positions = positions;
  a.rcov = m->rcov;
  a.r4r2 = m->r4r2;
  a.rcov_per_element = m->rcov_per_element;
  a.r4r2_per_element = m->r4r2_per_element;
  a.refcn = m->refcn;
  a.refc6 = m->refc6;
  a.rvdw = m->rvdw;
  a.rvdw_mode = m->rvdw_mode;
  a.order = m->batch_order;
  auto mk = [&](double v, double d) { return d3::make<S, B>((B)v, (B)d); };
  a.s6 = mk(par->s6, dpar ? dpar->s6 : 0.0);
  a.s8 = mk(par->s8, dpar ? dpar->s8 : 0.0);
  a.a1 = mk(par->a1, dpar ? dpar->a1 : 0.0);
  a.a2 = mk(par->a2, dpar ? dpar->a2 : 0.0);
  a.s9 = mk(par->s9, dpar ? dpar->s9 : 0.0);
  a.rs9 = mk(par->rs9, 0.0);
  a.alp = mk(par->alp, 0.0);
  // the reference compares distances in the working precision (disp.py:286)
  B cut = (B)par->cutoff, cncut = (B)par->cn_cutoff;
  a.cutoff2 = cut * cut;
  a.cn_cutoff2 = cncut * cncut;
  a.kcn = (B)par->kcn;
  if (par->counting < 0 || par->counting > 1 || par->damping < 0 || par->damping > 1)
    return fail(D3B200_EINVAL, "unknown counting / damping function id%s");
  a.counting = par->counting;
  a.damping = par->damping;
  return D3B200_OK;
}

// the pairs-once kernel (molecule_kernel_v2) implements the default model functions only
inline bool default_model(const d3b200_params* par) { return par->counting == 0 && par->damping == 0; }

template <typename S>
size_t workspace_bytes(int nbatch, int natoms, int atm) {
  if (!atm || nbatch <= 0 || natoms <= 0) return 0;
  return (size_t)nbatch * natoms * natoms * sizeof(S);
}

template <typename S>
int set_workspace(d3::MolArgs<S>& a, const d3b200_params* par, void* workspace, size_t bytes) {
  a.work = nullptr;
  if (par->s9 == 0.0) return D3B200_OK;
  size_t need = workspace_bytes<S>(a.nbatch, a.natoms, 1);
  if (need && (!workspace || bytes < need))
    return fail(D3B200_EINVAL, "ATM workspace too small%s: %ld < %ld bytes", "", (long)bytes, (long)need);
  a.work = reinterpret_cast<S*>(workspace);
  return D3B200_OK;
}

template <typename B>
int energy_impl(int nbatch, int natoms, const int64_t* numbers, const B* positions,
                const typename ModelOf<B>::type* model, const d3b200_params* par,
                B* energy, void* workspace, size_t wbytes, void* stream) {
  d3::MolArgs<B> a;
  int rc = fill_common<B, B>(a, nbatch, natoms, numbers, positions, model, par, nullptr);
  if (rc) return rc;
  if (!energy) return fail(D3B200_EINVAL, "null output%s");
  if (nbatch == 0 || natoms == 0) return D3B200_OK;
  a.energy = energy;
  if ((rc = set_workspace(a, par, workspace, wbytes))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (par->s9 != 0.0) return launch<B, d3::kWantE | d3::kAtm>(a, s);
  if (v2_fits<B>(natoms) && default_model(par)) return launch_v2<B, false>(a, s);
  return launch<B, d3::kWantE>(a, s);
}

template <typename B>
int vjp_impl(int nbatch, int natoms, const int64_t* numbers, const B* positions,
             const typename ModelOf<B>::type* model, const d3b200_params* par, const B* g,
             B* energy, B* grad, B* pgrad, void* workspace, size_t wbytes, void* stream, int host_io = 0,
             const int* ready = nullptr, int ready_per = 1) {
  d3::MolArgs<B> a;
  int rc = fill_common<B, B>(a, nbatch, natoms, numbers, positions, model, par, nullptr);
  if (rc) return rc;
  a.host_io = host_io;
  a.ready = ready;
  a.ready_per = ready_per;
  if (!grad) return fail(D3B200_EINVAL, "null gradient output%s");  // g == NULL: unit cotangent
  if (nbatch == 0 || natoms == 0) return D3B200_OK;
  a.g = g;
  a.energy = energy;
  a.grad = grad;
  a.pgrad = pgrad;
  if ((rc = set_workspace(a, par, workspace, wbytes))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const bool atm = par->s9 != 0.0;
  using namespace d3;
  if (pgrad) {
    if (atm) return launch<B, kWantE | kWantG | kWantP | kAtm>(a, s);
    return launch<B, kWantE | kWantG | kWantP>(a, s);
  }
  if (energy) {
    if (atm) return launch<B, kWantE | kWantG | kAtm>(a, s);
    if (v2_fits<B>(natoms) && default_model(par)) return launch_v2<B, true>(a, s);
    return launch<B, kWantE | kWantG>(a, s);
  }
  if (atm) return launch<B, kWantG | kAtm>(a, s);
  if (v2_fits<B>(natoms) && default_model(par)) return launch_v2<B, true>(a, s);
  return launch<B, kWantG>(a, s);
}

template <typename B>
int hvp_impl(int nbatch, int natoms, const int64_t* numbers, const B* positions,
             const B* dpositions, const typename ModelOf<B>::type* model,
             const d3b200_params* par, const d3b200_params* dpar, const B* g, B* grad,
             B* pgrad, B* denergy, B* dgrad, B* dpgrad, void* workspace, size_t wbytes,
             void* stream) {
  using S = d3::Dual<B>;
  d3::MolArgs<S> a;
  int rc = fill_common<S, B>(a, nbatch, natoms, numbers, positions, model, par, dpar);
  if (rc) return rc;
  if (!g) return fail(D3B200_EINVAL, "null cotangent%s");
  if (nbatch == 0 || natoms == 0) return D3B200_OK;
  a.dpositions = dpositions;
  a.g = g;
  a.grad = grad;
  a.pgrad = pgrad;
  a.denergy = denergy;
  a.dgrad = dgrad;
  a.dpgrad = dpgrad;
  if ((rc = set_workspace(a, par, workspace, wbytes))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  using namespace d3;
  if (par->s9 != 0.0) return launch<S, kWantE | kWantG | kWantP | kAtm>(a, s);
  return launch<S, kWantE | kWantG | kWantP>(a, s);
}

// ---------------------------------------------------------------- host-buffer pipeline
}  // namespace

constexpr int kPipeMaxChunks = 64;
struct d3b200_pipe {
  cudaStream_t in, out, comp[2];
  cudaEvent_t start, done, ev_in[kPipeMaxChunks], ev_k[kPipeMaxChunks];
  int device;
  int* h_one;   // page-locked constant 1: the streamed form raises its per-chunk flags with 4-byte H2D COPIES
                // (a small cudaMemsetAsync may be a kernel, which could not be scheduled next to the resident
                // CTAs that are waiting for exactly that flag)
};

namespace {

inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

// page-locked host memory that kernels of the current device can address (UVA)
bool device_visible(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost && a.devicePointer == p;
}

template <typename B>
size_t host_workspace_bytes(int nbatch, int natoms, int with_g, int atm) {
  if (nbatch <= 0 || natoms <= 0) return 0;
  const size_t na = (size_t)nbatch * natoms;
  size_t b = align256(na * sizeof(int64_t)) + align256(na * 3 * sizeof(B));   // numbers, positions
  if (with_g) b += align256(na * sizeof(B));
  b += align256(na * sizeof(B)) + align256(na * 3 * sizeof(B));                 // energy, grad
  b += align256(workspace_bytes<B>(nbatch, natoms, atm));
  return b;
}

// Chunk boundaries of the staged form: `chunk` structures per chunk (<= 0: one wave of CTAs of the
// fused kernel, 148 SMs x 4), quarter- and half-size chunks at both ends so that the un-overlapped
// first H2D copy and last D2H copy are short; at most kPipeMaxChunks chunks.  bounds[0..n], returns n.
int pipe_schedule(int nbatch, int chunk, int* bounds) {
  if (chunk < 0) chunk = -chunk;
  if (chunk == 0) chunk = 592;
  if ((nbatch + chunk - 1) / chunk + 4 > kPipeMaxChunks) chunk = (nbatch + kPipeMaxChunks - 5) / (kPipeMaxChunks - 4);
  int nchunk = 0;
  const int q = chunk / 4, h = chunk / 2;
  int lo = 0, hi = nbatch;
  int head[2], tail[2], nh = 0, nt = 0;
  if (q > 0 && nbatch >= 2 * chunk) { head[nh++] = q; head[nh++] = h; tail[nt++] = q; tail[nt++] = h; }
  bounds[nchunk++] = 0;
  for (int k = 0; k < nh; ++k) { lo += head[k]; bounds[nchunk++] = lo; }
  for (int k = 0; k < nt; ++k) hi -= tail[k];
  while (lo + chunk < hi) { lo += chunk; bounds[nchunk++] = lo; }
  if (lo < hi) bounds[nchunk++] = hi;
  for (int k = nt - 1; k >= 0; --k) { hi += tail[k]; bounds[nchunk++] = hi; }
  return nchunk - 1;
}

template <typename B>
int vjp_host_impl(d3b200_pipe* pipe, int nbatch, int natoms, const int64_t* h_numbers, const B* h_positions,
                  const typename ModelOf<B>::type* model, const d3b200_params* par, const B* h_g,
                  B* h_energy, B* h_grad, void* workspace, size_t wbytes, int chunk, void* stream) {
  if (!pipe) return fail(D3B200_EINVAL, "null pipe%s");
  if (nbatch < 0 || natoms < 0) return fail(D3B200_EINVAL, "negative size%s");
  if (!h_numbers || !h_positions || !model || !par || !h_grad) return fail(D3B200_EINVAL, "null argument%s");
  if (!model->rcov_per_element || !model->r4r2_per_element || model->rvdw_mode == 2)
    return fail(D3B200_EINVAL, "host entry takes per-element rcov / r4r2 / rvdw tables only%s");
  if (nbatch == 0 || natoms == 0) return D3B200_OK;
  const int atm = par->s9 != 0.0;
  // Zero-copy form (chunk <= 0, two-body + CN, all buffers page-locked and mapped): ONE launch of the
  // fused kernel on the host pointers.  Every CTA reads its structure over PCIe at its start and writes
  // its results at its end -- all accesses of the kernel are coalesced for that purpose -- so the transfers
  // of the 592 resident CTAs overlap the arithmetic of the others without any staging or chunking.
  const bool out_mapped = device_visible(h_grad) && (!h_energy || device_visible(h_energy));
  if (chunk == 0 && !atm && v2_fits<B>(natoms) && default_model(par) && device_visible(h_numbers) &&
      device_visible(h_positions) && out_mapped && (!h_g || device_visible(h_g)))
    return vjp_impl<B>(nbatch, natoms, h_numbers, h_positions, model, par, h_g, h_energy, h_grad, nullptr,
                       nullptr, 0, stream, 1);
  // Hybrid form (chunk < 0: |chunk| structures per chunk; outputs page-locked and mapped): the inputs are staged
  // with the copy engine, chunk by chunk, while the kernels of the previous chunks run and WRITE THEIR RESULTS
  // DIRECTLY INTO HOST MEMORY (posted PCIe writes, coalesced blocks) -- no device->host copies at all.
  const bool hybrid = chunk < 0 && !atm && v2_fits<B>(natoms) && default_model(par) && out_mapped;
  if (chunk < 0) chunk = -chunk;
  // Streamed form (chunk in -15..-2: 2^|chunk| structures per chunk; outputs mapped): ONE launch of the fused kernel on
  // `stream`; its CTAs wait for per-chunk flags that the copy stream sets behind each chunk of inputs, and write
  // their results directly into host memory.  Copy engine and kernel overlap without any launch boundary.
  const bool streamed = hybrid && chunk >= 2 && chunk <= 15;
  const size_t need = host_workspace_bytes<B>(nbatch, natoms, h_g != nullptr, atm);
  if (!workspace || wbytes < need)
    return fail(D3B200_EINVAL, "host-entry workspace too small%s: %ld < %ld bytes", "", (long)wbytes, (long)need);
  int bounds[kPipeMaxChunks + 1];
  const int nchunk = pipe_schedule(nbatch, chunk, bounds);

  const size_t na = (size_t)nbatch * natoms;
  char* w = reinterpret_cast<char*>(workspace);
  int64_t* d_z = reinterpret_cast<int64_t*>(w); w += align256(na * sizeof(int64_t));
  B* d_x = reinterpret_cast<B*>(w); w += align256(na * 3 * sizeof(B));
  B* d_g = nullptr;
  if (h_g) { d_g = reinterpret_cast<B*>(w); w += align256(na * sizeof(B)); }
  B* d_e = reinterpret_cast<B*>(w); w += align256(na * sizeof(B));
  B* d_gr = reinterpret_cast<B*>(w); w += align256(na * 3 * sizeof(B));
  char* d_atm = w;
  const size_t atm_per = workspace_bytes<B>(1, natoms, atm);

  typename ModelOf<B>::type chunk_model = *model;
  chunk_model.batch_order = nullptr;                 // the hint refers to the whole batch, not to a chunk
  cudaError_t e;
#define D3_CK(call, what) if ((e = (call)) != cudaSuccess) return cuda_fail(e, what)
  if (streamed) {
    const int per = 1 << chunk;
    const int nflag = (nbatch + per - 1) / per;
    // flags live in the (otherwise unused) device energy staging area
    int* d_ready = reinterpret_cast<int*>(d_e);
    if ((size_t)nflag * sizeof(int) > na * sizeof(B)) return fail(D3B200_EINVAL, "streamed host form: too many chunks%s");
    D3_CK(cudaMemsetAsync(d_ready, 0, (size_t)nflag * sizeof(int), (cudaStream_t)stream), "cudaMemsetAsync");
    D3_CK(cudaEventRecord(pipe->start, (cudaStream_t)stream), "cudaEventRecord");
    D3_CK(cudaStreamWaitEvent(pipe->in, pipe->start, 0), "cudaStreamWaitEvent");
    // (D3B200_STREAMED_DEVICE_OUT=1: timing experiment only -- results stay in the device staging area)
    static const bool dev_out = getenv("D3B200_STREAMED_DEVICE_OUT") != nullptr;
    int rc = vjp_impl<B>(nbatch, natoms, d_z, d_x, &chunk_model, par, h_g ? d_g : nullptr,
                         dev_out ? nullptr : h_energy, dev_out ? d_gr : h_grad, nullptr,
                         nullptr, 0, stream, 1, d_ready, per);
    if (rc) return rc;
    // numbers (and the cotangent) travel on a second copy stream, next to the positions: every copy pays ~10 us of
    // DMA set-up, which in ONE stream made a chunk's transfer slower than its arithmetic
    D3_CK(cudaStreamWaitEvent(pipe->out, pipe->start, 0), "cudaStreamWaitEvent");
    const int nev = nflag < kPipeMaxChunks ? nflag : kPipeMaxChunks;
    for (int k = 0; k < nflag; ++k) {
      const int lo = k * per, nb = (lo + per <= nbatch ? per : nbatch - lo);
      const size_t o = (size_t)lo * natoms, cnt = (size_t)nb * natoms;
      D3_CK(cudaMemcpyAsync(d_z + o, h_numbers + o, cnt * sizeof(int64_t), cudaMemcpyHostToDevice, pipe->out), "H2D numbers");
      if (h_g) D3_CK(cudaMemcpyAsync(d_g + o, h_g + o, cnt * sizeof(B), cudaMemcpyHostToDevice, pipe->out), "H2D cotangent");
      D3_CK(cudaEventRecord(pipe->ev_in[k % nev], pipe->out), "cudaEventRecord");
      D3_CK(cudaMemcpyAsync(d_x + 3 * o, h_positions + 3 * o, cnt * 3 * sizeof(B), cudaMemcpyHostToDevice, pipe->in), "H2D positions");
      D3_CK(cudaStreamWaitEvent(pipe->in, pipe->ev_in[k % nev], 0), "cudaStreamWaitEvent");
      D3_CK(cudaMemcpyAsync(d_ready + k, pipe->h_one, sizeof(int), cudaMemcpyHostToDevice, pipe->in), "H2D flag");
    }
    // the next call may reuse the staging buffers: `stream` also waits for the copy stream (already implied by the
    // kernel having consumed every flag, made explicit for the stream order)
    D3_CK(cudaEventRecord(pipe->done, pipe->in), "cudaEventRecord");
    D3_CK(cudaStreamWaitEvent((cudaStream_t)stream, pipe->done, 0), "cudaStreamWaitEvent");
    return D3B200_OK;
  }
  D3_CK(cudaEventRecord(pipe->start, (cudaStream_t)stream), "cudaEventRecord");
  D3_CK(cudaStreamWaitEvent(pipe->in, pipe->start, 0), "cudaStreamWaitEvent");
  for (int k = 0; k < nchunk; ++k) {
    const int lo = bounds[k], nb = bounds[k + 1] - lo;
    const size_t o = (size_t)lo * natoms, cnt = (size_t)nb * natoms;
    D3_CK(cudaMemcpyAsync(d_z + o, h_numbers + o, cnt * sizeof(int64_t), cudaMemcpyHostToDevice, pipe->in), "H2D numbers");
    D3_CK(cudaMemcpyAsync(d_x + 3 * o, h_positions + 3 * o, cnt * 3 * sizeof(B), cudaMemcpyHostToDevice, pipe->in), "H2D positions");
    if (h_g) D3_CK(cudaMemcpyAsync(d_g + o, h_g + o, cnt * sizeof(B), cudaMemcpyHostToDevice, pipe->in), "H2D cotangent");
    D3_CK(cudaEventRecord(pipe->ev_in[k], pipe->in), "cudaEventRecord");
    cudaStream_t cs = pipe->comp[k & 1];
    D3_CK(cudaStreamWaitEvent(cs, pipe->ev_in[k], 0), "cudaStreamWaitEvent");
    int rc;
    if (hybrid)
      rc = vjp_impl<B>(nb, natoms, d_z + o, d_x + 3 * o, &chunk_model, par, h_g ? d_g + o : nullptr,
                       h_energy ? h_energy + o : nullptr, h_grad + 3 * o, nullptr, nullptr, 0, cs, 1);
    else
      rc = vjp_impl<B>(nb, natoms, d_z + o, d_x + 3 * o, &chunk_model, par, h_g ? d_g + o : nullptr,
                       h_energy ? d_e + o : nullptr, d_gr + 3 * o, nullptr,
                       atm ? d_atm + (size_t)lo * atm_per : nullptr, (size_t)nb * atm_per, cs);
    if (rc) return rc;
    D3_CK(cudaEventRecord(pipe->ev_k[k], cs), "cudaEventRecord");
    D3_CK(cudaStreamWaitEvent(pipe->out, pipe->ev_k[k], 0), "cudaStreamWaitEvent");
    if (hybrid) continue;
    if (h_energy) D3_CK(cudaMemcpyAsync(h_energy + o, d_e + o, cnt * sizeof(B), cudaMemcpyDeviceToHost, pipe->out), "D2H energy");
    D3_CK(cudaMemcpyAsync(h_grad + 3 * o, d_gr + 3 * o, cnt * 3 * sizeof(B), cudaMemcpyDeviceToHost, pipe->out), "D2H gradient");
  }
  D3_CK(cudaEventRecord(pipe->done, pipe->out), "cudaEventRecord");
  D3_CK(cudaStreamWaitEvent((cudaStream_t)stream, pipe->done, 0), "cudaStreamWaitEvent");
#undef D3_CK
  return D3B200_OK;
}

// ---------------------------------------------------------------- tiled system
template <typename B> struct SysOf;
template <> struct SysOf<double> { using type = d3b200_system_f64; };
template <> struct SysOf<float> { using type = d3b200_system_f32; };

template <typename B>
int sys_fill(d3::SysArgs<B>& a, const typename SysOf<B>::type* s, const d3b200_params* par, int align = d3::kSysTile) {
  if (!s || !par) return fail(D3B200_EINVAL, "null argument%s");
  if (s->natoms <= 0 || s->natoms % d3::kSysTile) return fail(D3B200_EINVAL, "natoms must be a positive multiple of 128%s");
  if (s->row_begin < 0 || s->row_end > s->natoms || s->row_begin > s->row_end ||
      s->row_begin % align || s->row_end % align)
    return fail(D3B200_EINVAL, "row range must be tile aligned%s");
  if (!s->atoms || !s->z || !s->box) return fail(D3B200_EINVAL, "null system array%s");
  memset(&a, 0, sizeof(a));
  a.natoms = s->natoms; a.row_begin = s->row_begin; a.row_end = s->row_end;
  a.atoms = reinterpret_cast<const d3::Atom4<B>*>(s->atoms);
  a.aux = reinterpret_cast<const d3::Pair2<B>*>(s->aux);
  a.z = s->z; a.box = s->box; a.w = s->w; a.dw = s->dw; a.wout = s->w; a.dwout = s->dw;
  a.refcn = s->refcn; a.refc6 = s->refc6;
  a.cn = s->cn; a.decn = s->decn; a.energy = s->energy; a.grad = s->grad;
  a.jsplit = s->jsplit; a.part = s->scratch;
  a.nelem = 0; a.zlist = nullptr; a.eidx = nullptr; a.tvec = nullptr;
  if (s->nelem > 0 && s->nelem <= d3::kSysEmax && s->zlist && s->eidx && s->tvec) {
    a.nelem = s->nelem; a.zlist = s->zlist; a.eidx = reinterpret_cast<const signed char*>(s->eidx); a.tvec = s->tvec;
  }
  if (s->jsplit < 1 || !s->scratch) return fail(D3B200_EINVAL, "system scratch / jsplit missing%s");
  a.s6 = (B)par->s6; a.s8 = (B)par->s8; a.a1 = (B)par->a1; a.a2 = (B)par->a2;
  B cut = (B)par->cutoff, cncut = (B)par->cn_cutoff;
  a.cutoff2 = cut * cut; a.cn_cutoff2 = cncut * cncut; a.kcn = (B)par->kcn;
  return D3B200_OK;
}

int launched(const char* what) {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? D3B200_OK : cuda_fail(e, what);
}

// The 8-row-group CN sweeps pay while a CTA walks many j-tiles (one GPU: 0.61 -> 0.50 ms on the 50k-atom cluster); with
// the fine j-split of 8 ranks (73 CTAs per row tile, a few j-tiles each) their per-CTA set-up costs more than the
// better lane efficiency gains (0.138 -> 0.160 ms), so the 32-row form is kept there.
inline bool sym_use_q8(int jsplit) { return D3_SYM_Q8 != 0 && jsplit <= 40; }

// row tiles owned by a symmetric launch
template <typename T>
int sym_tiles(const T* s) {
  const int stride = s->tile_stride > 0 ? s->tile_stride : 1;
  const int span = (s->row_end - s->row_begin) / d3::kSysTile;
  return (span + stride - 1) / stride;
}

template <typename B>
int sys_cn(const typename SysOf<B>::type* s, const d3b200_params* par, void* stream) {
  d3::SysArgs<B> a;
  int rc = sys_fill<B>(a, s, par);
  if (rc) return rc;
  if (!s->cn) return fail(D3B200_EINVAL, "null cn%s");
  const int nblk = (a.row_end - a.row_begin) / d3::kSysTile;
  if (nblk == 0) return D3B200_OK;
  if (s->symmetric) {
    d3::SymArgs<B> p{a, s->tile_stride > 0 ? s->tile_stride : 1};
    if (sym_use_q8(a.jsplit)) d3::sym_cn_q8_kernel<B><<<dim3(a.jsplit, sym_tiles(s)), d3::kSysTile, 0, (cudaStream_t)stream>>>(p);
    else d3::sym_cn_kernel<B><<<dim3(a.jsplit, sym_tiles(s)), d3::kSysTile, 0, (cudaStream_t)stream>>>(p);
    return launched("sym_cn_kernel");
  }
  d3::system_cn_kernel<B><<<dim3(nblk, a.jsplit), d3::kSysTile, 0, (cudaStream_t)stream>>>(a);
  if ((rc = launched("system_cn_kernel"))) return rc;
  d3::system_reduce_kernel<B, 0><<<nblk, d3::kSysTile, 0, (cudaStream_t)stream>>>(a);
  return launched("system_reduce_kernel");
}

template <typename B>
int sys_weights(const typename SysOf<B>::type* s, const d3b200_params* par, void* stream) {
  d3::SysArgs<B> a;
  int rc = sys_fill<B>(a, s, par);
  if (rc) return rc;
  if (!s->cn || !s->w || !s->dw || !s->refcn) return fail(D3B200_EINVAL, "null weights argument%s");
  d3::system_weights_kernel<B><<<(a.natoms + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
  return launched("system_weights_kernel");
}

template <typename B>
int sys_tvec(const typename SysOf<B>::type* s, const d3b200_params* par, void* stream) {
  d3::SysArgs<B> a;
  int rc = sys_fill<B>(a, s, par);
  if (rc) return rc;
  if (!a.nelem || !s->w || !s->dw || !s->refc6) return fail(D3B200_EINVAL, "tvec needs nelem/zlist/eidx/tvec, w and refc6%s");
  const int nt = a.natoms * a.nelem;
  d3::system_tvec_kernel<B><<<(nt + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a.natoms, a.nelem, a.z, a.zlist, a.w, a.dw, a.refc6, a.tvec);
  return launched("system_tvec_kernel");
}

template <typename B>
int sys_pair(const typename SysOf<B>::type* s, const d3b200_params* par, int want_grad, void* stream) {
  d3::SysArgs<B> a;
  int rc = sys_fill<B>(a, s, par);
  if (rc) return rc;
  if (!s->aux || !s->w || !s->dw || !s->refc6) return fail(D3B200_EINVAL, "null pair argument%s");
  if (want_grad && (!s->grad || !s->decn)) return fail(D3B200_EINVAL, "null gradient output%s");
  const int nblk = (a.row_end - a.row_begin) / d3::kSysTile;
  if (nblk == 0) return D3B200_OK;
  if (!want_grad) a.grad = nullptr;
  if (s->symmetric) {
    if (!a.nelem) return fail(D3B200_EINVAL, "symmetric pair sweep needs the T vectors (nelem <= 8)%s");
    if (!s->energy) return fail(D3B200_EINVAL, "symmetric pair sweep needs the energy array%s");
    d3::SymArgs<B> p{a, s->tile_stride > 0 ? s->tile_stride : 1};
    const size_t smem = d3::sym_pair_smem_bytes<B>(a.nelem);
    const dim3 grid(a.jsplit, sym_tiles(s));
#if D3_SYM_Q8 && D3_SYM_PAIR_Q8
    if (want_grad) d3::sym_pair_q8_kernel<B, true><<<grid, d3::kSysTile, 0, (cudaStream_t)stream>>>(p);
    else d3::sym_pair_q8_kernel<B, false><<<grid, d3::kSysTile, 0, (cudaStream_t)stream>>>(p);
    return launched("sym_pair_q8_kernel");
#endif
    if (want_grad) {
      auto kern = d3::sym_pair_t_kernel<B, true>;
      if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<grid, d3::kSysTile, smem, (cudaStream_t)stream>>>(p);
    } else {
      auto kern = d3::sym_pair_t_kernel<B, false>;
      if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<grid, d3::kSysTile, smem, (cudaStream_t)stream>>>(p);
    }
    return launched("sym_pair_t_kernel");
  }
  if (a.nelem) {
    const size_t smem = d3::pair_t_smem_bytes<B>(a.nelem);
    if (want_grad) {
      auto kern = d3::system_pair_t_kernel<B, true>;
      if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<dim3(nblk, a.jsplit), d3::kSysTile, smem, (cudaStream_t)stream>>>(a);
    } else {
      auto kern = d3::system_pair_t_kernel<B, false>;
      if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<dim3(nblk, a.jsplit), d3::kSysTile, smem, (cudaStream_t)stream>>>(a);
    }
  } else if (want_grad) d3::system_pair_kernel<B, true><<<dim3(nblk, a.jsplit), d3::kSysTile, 0, (cudaStream_t)stream>>>(a);
  else d3::system_pair_kernel<B, false><<<dim3(nblk, a.jsplit), d3::kSysTile, 0, (cudaStream_t)stream>>>(a);
  if ((rc = launched("system_pair_kernel"))) return rc;
  d3::system_reduce_kernel<B, 1><<<nblk, d3::kSysTile, 0, (cudaStream_t)stream>>>(a);
  return launched("system_reduce_kernel");
}

template <typename B>
int sys_pack(int natoms, int nreal, const int64_t* perm, const B* positions, const B* tmpl, const int32_t* z,
             B* atoms, B* box, void* stream) {
  if (natoms <= 0 || natoms % d3::kSysTile || nreal < 0 || nreal > natoms)
    return fail(D3B200_EINVAL, "system_pack: natoms must be a positive multiple of 128 and 0 <= nreal <= natoms%s");
  if ((nreal && (!perm || !positions)) || !tmpl || !z || !atoms || !box) return fail(D3B200_EINVAL, "system_pack: null argument%s");
  const int nsub = natoms / d3::kSysSub;
  d3::system_pack_kernel<B><<<(nsub + 3) / 4, 128, 0, (cudaStream_t)stream>>>(natoms, nreal, perm, positions, tmpl, z, atoms, box);
  return launched("system_pack_kernel");
}

template <typename B>
int sys_pgrad(const typename SysOf<B>::type* s, const d3b200_params* par, void* stream) {
  d3::SysArgs<B> a;
  int rc = sys_fill<B>(a, s, par);
  if (rc) return rc;
  if (!s->aux || !s->w) return fail(D3B200_EINVAL, "null pgrad argument%s");
  if (!a.nelem) return fail(D3B200_EINVAL, "the parameter-gradient sweep needs the T vectors (nelem <= 8)%s");
  const int nblk = (a.row_end - a.row_begin) / d3::kSysTile;
  if (nblk == 0) return D3B200_OK;
  const size_t smem = d3::pair_t_smem_bytes<B>(a.nelem);
  auto kern = d3::system_pgrad_kernel<B>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<dim3(nblk, a.jsplit), d3::kSysTile, smem, (cudaStream_t)stream>>>(a);
  return launched("system_pgrad_kernel");
}

template <typename B>
int sys_cnbwd(const typename SysOf<B>::type* s, const d3b200_params* par, void* stream) {
  d3::SysArgs<B> a;
  int rc = sys_fill<B>(a, s, par);
  if (rc) return rc;
  if (!s->decn || !s->grad) return fail(D3B200_EINVAL, "null cnbwd argument%s");
  const int nblk = (a.row_end - a.row_begin) / d3::kSysTile;
  if (nblk == 0) return D3B200_OK;
  if (s->symmetric) {
    d3::SymArgs<B> p{a, s->tile_stride > 0 ? s->tile_stride : 1};
    if (sym_use_q8(a.jsplit)) d3::sym_cnbwd_q8_kernel<B><<<dim3(a.jsplit, sym_tiles(s)), d3::kSysTile, 0, (cudaStream_t)stream>>>(p);
    else d3::sym_cnbwd_kernel<B><<<dim3(a.jsplit, sym_tiles(s)), d3::kSysTile, 0, (cudaStream_t)stream>>>(p);
    return launched("sym_cnbwd_kernel");
  }
  d3::system_cnbwd_kernel<B><<<dim3(nblk, a.jsplit), d3::kSysTile, 0, (cudaStream_t)stream>>>(a);
  if ((rc = launched("system_cnbwd_kernel"))) return rc;
  d3::system_reduce_kernel<B, 2><<<nblk, d3::kSysTile, 0, (cudaStream_t)stream>>>(a);
  return launched("system_reduce_kernel");
}

// CTAs of the persistent ATM kernel: resident CTAs per SM x SMs of the current device
int atm3_grid() {
  int dev = 0, nsm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0) {
    cudaGetLastError();
    nsm = 148;
  }
  return nsm * d3::kAtm3Ctas;
}

// Largest pair table (bytes) the triples-once ATM kernel may place in `atm_work`: 32 natoms^2 bytes in fp64
// (3.2 GB for 10 k atoms).  Beyond it the sweep contracts C6_jk on the fly.  $D3B200_ATM_PAIRTAB_MAX_BYTES, 0 = never.
size_t atm3_pairtab_limit() {
  const char* e = getenv("D3B200_ATM_PAIRTAB_MAX_BYTES");
  if (e && *e) return (size_t)strtoull(e, nullptr, 10);
  return (size_t)8 << 30;
}
template <typename B>
size_t atm3_total_workspace_bytes(int natoms) {
  const size_t scratch = d3::atm3_workspace_bytes<B>(natoms, atm3_grid());
  const size_t tab = d3::atm3_pairtab_bytes<B>(natoms);
  return tab <= atm3_pairtab_limit() ? scratch + tab : scratch;
}

template <typename B>
int sys_atm(const typename SysOf<B>::type* s, const d3b200_params* par, const B* rvdw, int want_grad, void* stream) {
  d3::AtmArgs<B> p;
  // tile_stride < 0: centres dealt atom by atom (row_begin = first centre, any value), see d3b200.h
  int rc = sys_fill<B>(p.sys, s, par, s && s->tile_stride < 0 ? 1 : d3::kSysSub);
  if (rc) return rc;
  if (!rvdw || !s->aux || !s->w || !s->dw || !s->refc6 || !s->energy) return fail(D3B200_EINVAL, "null ATM argument%s");
  if (want_grad && (!s->grad || !s->decn)) return fail(D3B200_EINVAL, "null ATM gradient output%s");
  p.rvdw = rvdw;
  p.s9 = (B)par->s9; p.rs9 = (B)par->rs9; p.pexp = (B)((par->alp + 2.0) / 3.0);
  const int rows = p.sys.row_end - p.sys.row_begin;
  if (rows == 0 || par->s9 == 0.0) return D3B200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (p.sys.nelem) {
    d3::AtmArgs2<B> q;
    q.sys = p.sys;
    q.nelem = p.sys.nelem;
    q.eidx = p.sys.eidx;
    q.tvec = p.sys.tvec;
    B* rve = p.sys.tvec + (size_t)p.sys.natoms * q.nelem * 16;  // 64 spare reals behind T and Td
    q.rvdw_e = rve;
    q.s9 = p.s9; q.rs9 = p.rs9; q.pexp = p.pexp;
    q.cube_root_pow = par->alp == 14.0 ? 1 : 0;
    d3::system_rvdw_compact_kernel<B><<<1, 64, 0, st>>>(q.nelem, p.sys.zlist, rvdw, rve);
    if (s->atm_work && s->atm_work_bytes >= d3::atm3_workspace_bytes<B>(p.sys.natoms, atm3_grid())) {
      // every triple once: persistent CTAs, centre atoms from a counter
      d3::AtmArgs3<B> r;
      r.a = q;
      r.tile_stride = s->tile_stride != 0 ? s->tile_stride : 1;
      const int grid = atm3_grid();
      {
        // about four work units per persistent CTA, at most 8 per centre (eight per CTA measured slower:
        // 10 k atoms on 8 GPUs 12.8 ms against 10.8 ms, profiles/r2_summary.md)
        const int astride = r.tile_stride < 0 ? -r.tile_stride : r.tile_stride;
        const int owned = (rows + astride - 1) / astride;
        int parts = owned > 0 ? (4 * grid + owned - 1) / owned : 1;
        r.parts = parts < 1 ? 1 : (parts > 8 ? 8 : parts);
      }
      char* wsp = reinterpret_cast<char*>(s->atm_work);
      r.counter = reinterpret_cast<int*>(wsp);
      r.rec = reinterpret_cast<B*>(wsp + 256);
      r.list = reinterpret_cast<int*>(wsp + 256 + (size_t)grid * p.sys.natoms * d3::kAtm3Rec * sizeof(B));
      cudaError_t e = cudaMemsetAsync(r.counter, 0, 2 * sizeof(int), st);
      if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
      // largest neighbour count -> stride of the per-CTA neighbour slots (keeps the scratch L2-resident)
      d3::system_atm_maxn_kernel<B><<<p.sys.natoms / d3::kSysSub, 32 * d3::kMaxnWarps, 0, st>>>(p.sys, q.eidx, r.counter + 1);
      const size_t scratch = d3::atm3_workspace_bytes<B>(p.sys.natoms, grid);
      const size_t tabb = d3::atm3_pairtab_bytes<B>(p.sys.natoms);
      r.tab = nullptr;
      if (tabb <= atm3_pairtab_limit() && s->atm_work_bytes >= scratch + tabb) {
        // everything that belongs to an edge jk alone, once per pair (upper triangle of a dense table)
        B* tab = reinterpret_cast<B*>(wsp + scratch);
        r.tab = tab;
        const dim3 tg((p.sys.natoms + d3::kPairTabCols - 1) / d3::kPairTabCols, p.sys.natoms / d3::kPairTabRows);
        d3::system_atm_pairtab_kernel<B><<<tg, d3::kPairTabCols, 0, st>>>(q, tab);
        if (want_grad) d3::system_atm3_kernel<B, true, true><<<grid, 32 * d3::kAtm3Warps, 0, st>>>(r);
        else d3::system_atm3_kernel<B, false, true><<<grid, 32 * d3::kAtm3Warps, 0, st>>>(r);
        return launched("system_atm3_kernel");
      }
      if (want_grad) d3::system_atm3_kernel<B, true><<<grid, 32 * d3::kAtm3Warps, 0, st>>>(r);
      else d3::system_atm3_kernel<B, false><<<grid, 32 * d3::kAtm3Warps, 0, st>>>(r);
      return launched("system_atm3_kernel");
    }
    if (s->tile_stride > 1 || s->tile_stride < 0) return fail(D3B200_EINVAL, "strided ATM ownership needs atm_work (the triples-once kernel)%s");
    const size_t smem = d3::atm2_smem_bytes<B>(q.nelem);
    if (want_grad) {
      auto kern = d3::system_atm2_kernel<B, true>;
      if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<rows, 32 * d3::kAtmWarps, smem, st>>>(q);
    } else {
      auto kern = d3::system_atm2_kernel<B, false>;
      if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<rows, 32 * d3::kAtmWarps, smem, st>>>(q);
    }
    return launched("system_atm2_kernel");
  }
  if (s->tile_stride > 1 || s->tile_stride < 0) return fail(D3B200_EINVAL, "strided ATM ownership needs the T vectors and atm_work%s");
  if (want_grad) d3::system_atm_kernel<B, true><<<rows, 32 * d3::kAtmWarps, 0, st>>>(p);
  else d3::system_atm_kernel<B, false><<<rows, 32 * d3::kAtmWarps, 0, st>>>(p);
  return launched("system_atm_kernel");
}

// ---------------------------------------------------------------- stages
template <typename B>
int stage_cn(int nb, int n, const int64_t* z, const B* x, const B* rcov, double cutoff, double kcn, B* cn, void* st) {
  if (nb < 0 || n < 0 || !z || !x || !rcov || !cn) return fail(D3B200_EINVAL, "bad stage_cn argument%s");
  if (!nb || !n) return D3B200_OK;
  B c = (B)cutoff;
  d3::stage_cn_kernel<B><<<dim3((n + 127) / 128, nb), 128, 0, (cudaStream_t)st>>>(n, z, x, rcov, c * c, (B)kcn, cn);
  return launched("stage_cn_kernel");
}
template <typename B>
int stage_cn_model(int nb, int n, const int64_t* z, const B* x, const B* rcov, double cutoff, double kcn, int counting,
                   B* cn, void* st) {
  if (nb < 0 || n < 0 || !z || !x || !rcov || !cn || counting < 0 || counting > 1) return fail(D3B200_EINVAL, "bad stage_cn_model argument%s");
  if (!nb || !n) return D3B200_OK;
  B c = (B)cutoff;
  d3::stage_cn_kernel<B><<<dim3((n + 127) / 128, nb), 128, 0, (cudaStream_t)st>>>(n, z, x, rcov, c * c, (B)kcn, cn, counting);
  return launched("stage_cn_kernel");
}
template <typename B>
int stage_cn_model_bwd(int nb, int n, const int64_t* z, const B* x, const B* rcov, double cutoff, double kcn, int counting,
                       const B* gcn, B* gx, void* st) {
  if (nb < 0 || n < 0 || !z || !x || !rcov || !gcn || !gx || counting < 0 || counting > 1) return fail(D3B200_EINVAL, "bad stage_cn_model_bwd argument%s");
  if (!nb || !n) return D3B200_OK;
  B c = (B)cutoff;
  d3::stage_cn_bwd_kernel<B><<<dim3((n + 127) / 128, nb), 128, 0, (cudaStream_t)st>>>(n, z, x, rcov, c * c, (B)kcn, gcn, gx, counting);
  return launched("stage_cn_bwd_kernel");
}
template <typename B>
int stage_weights(size_t n, const int64_t* z, const B* cn, const B* refcn, B* w, void* st) {
  if (!z || !cn || !refcn || !w) return fail(D3B200_EINVAL, "bad stage_weights argument%s");
  if (!n) return D3B200_OK;
  d3::stage_weights_kernel<B><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)st>>>(n, z, cn, refcn, w);
  return launched("stage_weights_kernel");
}
template <typename B>
int stage_c6(int nb, int n, const int64_t* z, const B* w, const B* refc6, B* c6, void* st) {
  if (nb < 0 || n < 0 || !z || !w || !refc6 || !c6) return fail(D3B200_EINVAL, "bad stage_c6 argument%s");
  if (!nb || !n) return D3B200_OK;
  if (nb > 65535 || n > 65535) return fail(D3B200_ESIZE, "stage_c6: batch or structure too large%s");
  d3::stage_c6_kernel<B><<<dim3((n + 127) / 128, n, nb), 128, 0, (cudaStream_t)st>>>(n, z, w, refc6, c6);
  return launched("stage_c6_kernel");
}
template <typename B>
int stage_c6_bilinear(int nb, int n, const int64_t* z, const B* u, const B* w, const B* refc6, B* out, void* st) {
  if (nb < 0 || n < 0 || !z || !u || !w || !refc6 || !out) return fail(D3B200_EINVAL, "bad stage_c6_bilinear argument%s");
  if (!nb || !n) return D3B200_OK;
  if (nb > 65535 || n > 65535) return fail(D3B200_ESIZE, "stage_c6_bilinear: batch or structure too large%s");
  d3::stage_c6_bilinear_kernel<B><<<dim3((n + 127) / 128, n, nb), 128, 0, (cudaStream_t)st>>>(n, z, u, w, refc6, out);
  return launched("stage_c6_bilinear_kernel");
}
template <typename B>
int stage_c6_contract(int nb, int n, const int64_t* z, const B* G, const B* w, const B* refc6, B* out, void* st) {
  if (nb < 0 || n < 0 || !z || !G || !w || !refc6 || !out) return fail(D3B200_EINVAL, "bad stage_c6_contract argument%s");
  if (!nb || !n) return D3B200_OK;
  if (nb > 65535) return fail(D3B200_ESIZE, "stage_c6_contract: batch too large%s");
  d3::stage_c6_contract_kernel<B><<<dim3((n + 63) / 64, nb), 64, 0, (cudaStream_t)st>>>(n, z, G, w, refc6, out);
  return launched("stage_c6_contract_kernel");
}
template <typename B>
int stage_disp2(int nb, int n, const int64_t* z, const B* x, const B* c6, const B* q, const d3b200_params* p, B* e, void* st) {
  if (nb < 0 || n < 0 || !z || !x || !c6 || !q || !p || !e) return fail(D3B200_EINVAL, "bad stage_disp2 argument%s");
  if (!nb || !n) return D3B200_OK;
  B c = (B)p->cutoff;
  d3::stage_disp2_kernel<B><<<dim3((n + 127) / 128, nb), 128, 0, (cudaStream_t)st>>>(
      n, z, x, c6, q, (B)p->s6, (B)p->s8, (B)p->a1, (B)p->a2, c * c, e, p->damping, (B)(p->alp > 0 ? p->alp : 14.0));
  return launched("stage_disp2_kernel");
}
template <typename B>
int stage_atm(int nb, int n, const int64_t* z, const B* x, const B* c6, const B* rvdw, const d3b200_params* p, B* e, void* st) {
  if (nb < 0 || n < 0 || !z || !x || !c6 || !rvdw || !p || !e) return fail(D3B200_EINVAL, "bad stage_atm argument%s");
  if (!nb || !n) return D3B200_OK;
  B c = (B)p->cutoff;
  d3::stage_atm_kernel<B><<<dim3((n + 63) / 64, nb), 64, 0, (cudaStream_t)st>>>(
      n, z, x, c6, rvdw, c * c, (B)p->s9, (B)p->rs9, (B)((p->alp + 2.0) / 3.0), e);
  return launched("stage_atm_kernel");
}

template <typename B>
int stage_cn_bwd(int nb, int n, const int64_t* z, const B* x, const B* rcov, double cutoff, double kcn, const B* gcn,
                 B* gx, void* st) {
  if (nb < 0 || n < 0 || !z || !x || !rcov || !gcn || !gx) return fail(D3B200_EINVAL, "bad stage_cn_bwd argument%s");
  if (!nb || !n) return D3B200_OK;
  B c = (B)cutoff;
  d3::stage_cn_bwd_kernel<B><<<dim3((n + 127) / 128, nb), 128, 0, (cudaStream_t)st>>>(n, z, x, rcov, c * c, (B)kcn, gcn, gx);
  return launched("stage_cn_bwd_kernel");
}
template <typename B>
int stage_weights_bwd(size_t n, const int64_t* z, const B* cn, const B* refcn, const B* gw, B* gcn, void* st) {
  if (!z || !cn || !refcn || !gw || !gcn) return fail(D3B200_EINVAL, "bad stage_weights_bwd argument%s");
  if (!n) return D3B200_OK;
  d3::stage_weights_bwd_kernel<B><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)st>>>(n, z, cn, refcn, gw, gcn);
  return launched("stage_weights_bwd_kernel");
}
template <typename B>
int stage_disp2_bwd(int nb, int n, const int64_t* z, const B* x, const B* c6, const B* q, const d3b200_params* p,
                    const B* ge, B* gx, B* gc6, B* gpar, void* st) {
  if (nb < 0 || n < 0 || !z || !x || !c6 || !q || !p || !ge || !gx || !gc6 || !gpar)
    return fail(D3B200_EINVAL, "bad stage_disp2_bwd argument%s");
  if (!nb || !n) return D3B200_OK;
  B c = (B)p->cutoff;
  d3::stage_disp2_bwd_kernel<B><<<dim3((n + 127) / 128, nb), 128, 0, (cudaStream_t)st>>>(
      n, z, x, c6, q, (B)p->s6, (B)p->s8, (B)p->a1, (B)p->a2, c * c, ge, gx, gc6, gpar, p->damping,
      (B)(p->alp > 0 ? p->alp : 14.0));
  return launched("stage_disp2_bwd_kernel");
}
template <typename B>
int stage_atm_bwd(int nb, int n, const int64_t* z, const B* x, const B* c6, const B* rvdw, const d3b200_params* p,
                  const B* ge, B* gx, B* gc6, void* st) {
  if (nb < 0 || n < 0 || !z || !x || !c6 || !rvdw || !p || !ge || !gx || !gc6)
    return fail(D3B200_EINVAL, "bad stage_atm_bwd argument%s");
  if (!nb || !n) return D3B200_OK;
  if (nb > 65535 || n > 65535) return fail(D3B200_ESIZE, "stage_atm_bwd: batch or structure too large%s");
  B c = (B)p->cutoff;
  d3::stage_atm_bwd_kernel<B><<<dim3((n + 63) / 64, n, nb), 64, 0, (cudaStream_t)st>>>(
      n, z, x, c6, rvdw, c * c, (B)p->s9, (B)p->rs9, (B)((p->alp + 2.0) / 3.0), ge, gx, gc6);
  return launched("stage_atm_bwd_kernel");
}

}  // namespace

extern "C" {

#define D3_STAGE_API(SFX, T)                                                                                  \
  int d3b200_stage_cn_##SFX(int nb, int n, const int64_t* z, const T* x, const T* rcov, double cutoff,       \
                            double kcn, T* cn, void* st) { return stage_cn<T>(nb, n, z, x, rcov, cutoff, kcn, cn, st); } \
  int d3b200_stage_cn_model_##SFX(int nb, int n, const int64_t* z, const T* x, const T* rcov, double cutoff, \
                                  double kcn, int counting, T* cn, void* st) {                               \
    return stage_cn_model<T>(nb, n, z, x, rcov, cutoff, kcn, counting, cn, st); }                             \
  int d3b200_stage_cn_model_bwd_##SFX(int nb, int n, const int64_t* z, const T* x, const T* rcov, double cutoff, \
                                      double kcn, int counting, const T* gcn, T* gx, void* st) {             \
    return stage_cn_model_bwd<T>(nb, n, z, x, rcov, cutoff, kcn, counting, gcn, gx, st); }                    \
  int d3b200_stage_weights_##SFX(size_t n, const int64_t* z, const T* cn, const T* refcn, T* w, void* st) {  \
    return stage_weights<T>(n, z, cn, refcn, w, st); }                                                        \
  int d3b200_stage_c6_##SFX(int nb, int n, const int64_t* z, const T* w, const T* refc6, T* c6, void* st) {  \
    return stage_c6<T>(nb, n, z, w, refc6, c6, st); }                                                         \
  int d3b200_stage_c6_bilinear_##SFX(int nb, int n, const int64_t* z, const T* u, const T* w, const T* refc6,   \
                                     T* out, void* st) { return stage_c6_bilinear<T>(nb, n, z, u, w, refc6, out, st); } \
  int d3b200_stage_c6_contract_##SFX(int nb, int n, const int64_t* z, const T* G, const T* w, const T* refc6,   \
                                     T* out, void* st) { return stage_c6_contract<T>(nb, n, z, G, w, refc6, out, st); } \
  int d3b200_stage_disp2_##SFX(int nb, int n, const int64_t* z, const T* x, const T* c6, const T* q,         \
                               const d3b200_params* p, T* e, void* st) { return stage_disp2<T>(nb, n, z, x, c6, q, p, e, st); } \
  int d3b200_stage_atm_##SFX(int nb, int n, const int64_t* z, const T* x, const T* c6, const T* rvdw,        \
                             const d3b200_params* p, T* e, void* st) { return stage_atm<T>(nb, n, z, x, c6, rvdw, p, e, st); } \
  int d3b200_stage_cn_bwd_##SFX(int nb, int n, const int64_t* z, const T* x, const T* rcov, double cutoff,   \
                                double kcn, const T* gcn, T* gx, void* st) {                                 \
    return stage_cn_bwd<T>(nb, n, z, x, rcov, cutoff, kcn, gcn, gx, st); }                                    \
  int d3b200_stage_weights_bwd_##SFX(size_t n, const int64_t* z, const T* cn, const T* refcn, const T* gw,   \
                                     T* gcn, void* st) { return stage_weights_bwd<T>(n, z, cn, refcn, gw, gcn, st); } \
  int d3b200_stage_disp2_bwd_##SFX(int nb, int n, const int64_t* z, const T* x, const T* c6, const T* q,     \
                                   const d3b200_params* p, const T* ge, T* gx, T* gc6, T* gpar, void* st) {  \
    return stage_disp2_bwd<T>(nb, n, z, x, c6, q, p, ge, gx, gc6, gpar, st); }                                \
  int d3b200_stage_atm_bwd_##SFX(int nb, int n, const int64_t* z, const T* x, const T* c6, const T* rvdw,    \
                                 const d3b200_params* p, const T* ge, T* gx, T* gc6, void* st) {             \
    return stage_atm_bwd<T>(nb, n, z, x, c6, rvdw, p, ge, gx, gc6, st); }
D3_STAGE_API(f64, double)
D3_STAGE_API(f32, float)
#undef D3_STAGE_API

size_t d3b200_system_atm_workspace_bytes_f64(int natoms) { return natoms > 0 ? atm3_total_workspace_bytes<double>(natoms) : 0; }
size_t d3b200_system_atm_workspace_bytes_f32(int natoms) { return natoms > 0 ? atm3_total_workspace_bytes<float>(natoms) : 0; }
int d3b200_system_atm_f64(const d3b200_system_f64* s, const d3b200_params* p, const double* r, int g, void* st) { return sys_atm<double>(s, p, r, g, st); }
int d3b200_system_atm_f32(const d3b200_system_f32* s, const d3b200_params* p, const float* r, int g, void* st) { return sys_atm<float>(s, p, r, g, st); }
int d3b200_system_tvec_f64(const d3b200_system_f64* s, const d3b200_params* p, void* st) { return sys_tvec<double>(s, p, st); }
int d3b200_system_tvec_f32(const d3b200_system_f32* s, const d3b200_params* p, void* st) { return sys_tvec<float>(s, p, st); }
int d3b200_system_pack_f64(int n, int nr, const int64_t* perm, const double* x, const double* t, const int32_t* z, double* a, double* b, void* st) { return sys_pack<double>(n, nr, perm, x, t, z, a, b, st); }
int d3b200_system_pack_f32(int n, int nr, const int64_t* perm, const float* x, const float* t, const int32_t* z, float* a, float* b, void* st) { return sys_pack<float>(n, nr, perm, x, t, z, a, b, st); }
int d3b200_system_pgrad_f64(const d3b200_system_f64* s, const d3b200_params* p, void* st) { return sys_pgrad<double>(s, p, st); }
int d3b200_system_pgrad_f32(const d3b200_system_f32* s, const d3b200_params* p, void* st) { return sys_pgrad<float>(s, p, st); }
int d3b200_system_cn_f64(const d3b200_system_f64* s, const d3b200_params* p, void* st) { return sys_cn<double>(s, p, st); }
int d3b200_system_weights_f64(const d3b200_system_f64* s, const d3b200_params* p, void* st) { return sys_weights<double>(s, p, st); }
int d3b200_system_pair_f64(const d3b200_system_f64* s, const d3b200_params* p, int g, void* st) { return sys_pair<double>(s, p, g, st); }
int d3b200_system_cnbwd_f64(const d3b200_system_f64* s, const d3b200_params* p, void* st) { return sys_cnbwd<double>(s, p, st); }
int d3b200_system_cn_f32(const d3b200_system_f32* s, const d3b200_params* p, void* st) { return sys_cn<float>(s, p, st); }
int d3b200_system_weights_f32(const d3b200_system_f32* s, const d3b200_params* p, void* st) { return sys_weights<float>(s, p, st); }
int d3b200_system_pair_f32(const d3b200_system_f32* s, const d3b200_params* p, int g, void* st) { return sys_pair<float>(s, p, g, st); }
int d3b200_system_cnbwd_f32(const d3b200_system_f32* s, const d3b200_params* p, void* st) { return sys_cnbwd<float>(s, p, st); }


int d3b200_abi_version(void) { return D3B200_ABI_VERSION; }
const char* d3b200_last_error(void) { return g_err; }

int d3b200_batch_max_atoms_f64(int order) {
  return order ? max_atoms<d3::Dual<double>>() : max_atoms<double>();
}
int d3b200_batch_max_atoms_f32(int order) {
  return order ? max_atoms<d3::Dual<float>>() : max_atoms<float>();
}

size_t d3b200_batch_workspace_bytes_f64(int nbatch, int natoms, int order, int atm) {
  return order ? workspace_bytes<d3::Dual<double>>(nbatch, natoms, atm)
               : workspace_bytes<double>(nbatch, natoms, atm);
}
size_t d3b200_batch_workspace_bytes_f32(int nbatch, int natoms, int order, int atm) {
  return order ? workspace_bytes<d3::Dual<float>>(nbatch, natoms, atm)
               : workspace_bytes<float>(nbatch, natoms, atm);
}

int d3b200_batch_energy_f64(int nbatch, int natoms, const int64_t* numbers,
                            const double* positions, const d3b200_model_f64* model,
                            const d3b200_params* par, double* energy, void* workspace,
                            size_t workspace_bytes, void* stream) {
  return energy_impl<double>(nbatch, natoms, numbers, positions, model, par, energy, workspace, workspace_bytes, stream);
}
int d3b200_batch_energy_f32(int nbatch, int natoms, const int64_t* numbers,
                            const float* positions, const d3b200_model_f32* model,
                            const d3b200_params* par, float* energy, void* workspace,
                            size_t workspace_bytes, void* stream) {
  return energy_impl<float>(nbatch, natoms, numbers, positions, model, par, energy, workspace, workspace_bytes, stream);
}

int d3b200_batch_vjp_f64(int nbatch, int natoms, const int64_t* numbers,
                         const double* positions, const d3b200_model_f64* model,
                         const d3b200_params* par, const double* g, double* energy,
                         double* grad, double* pgrad, void* workspace,
                         size_t workspace_bytes, void* stream) {
  return vjp_impl<double>(nbatch, natoms, numbers, positions, model, par, g, energy, grad, pgrad, workspace, workspace_bytes, stream);
}
int d3b200_batch_vjp_f32(int nbatch, int natoms, const int64_t* numbers,
                         const float* positions, const d3b200_model_f32* model,
                         const d3b200_params* par, const float* g, float* energy,
                         float* grad, float* pgrad, void* workspace,
                         size_t workspace_bytes, void* stream) {
  return vjp_impl<float>(nbatch, natoms, numbers, positions, model, par, g, energy, grad, pgrad, workspace, workspace_bytes, stream);
}

int d3b200_batch_hvp_f64(int nbatch, int natoms, const int64_t* numbers,
                         const double* positions, const double* dpositions,
                         const d3b200_model_f64* model, const d3b200_params* par,
                         const d3b200_params* dpar, const double* g, double* grad,
                         double* pgrad, double* denergy, double* dgrad,
                         double* dpgrad, void* workspace, size_t workspace_bytes,
                         void* stream) {
  return hvp_impl<double>(nbatch, natoms, numbers, positions, dpositions, model, par, dpar, g,
                          grad, pgrad, denergy, dgrad, dpgrad, workspace, workspace_bytes, stream);
}
int d3b200_batch_hvp_f32(int nbatch, int natoms, const int64_t* numbers,
                         const float* positions, const float* dpositions,
                         const d3b200_model_f32* model, const d3b200_params* par,
                         const d3b200_params* dpar, const float* g, float* grad,
                         float* pgrad, float* denergy, float* dgrad, float* dpgrad,
                         void* workspace, size_t workspace_bytes, void* stream) {
  return hvp_impl<float>(nbatch, natoms, numbers, positions, dpositions, model, par, dpar, g,
                         grad, pgrad, denergy, dgrad, dpgrad, workspace, workspace_bytes, stream);
}

int d3b200_pipe_create(d3b200_pipe** out) {
  if (!out) return fail(D3B200_EINVAL, "null argument%s");
  d3b200_pipe* p = (d3b200_pipe*)calloc(1, sizeof(d3b200_pipe));
  if (!p) return fail(D3B200_EINVAL, "out of host memory%s");
  cudaError_t e = cudaGetDevice(&p->device);
  cudaStream_t* ss[4] = {&p->in, &p->out, &p->comp[0], &p->comp[1]};
  for (int k = 0; k < 4 && e == cudaSuccess; ++k) e = cudaStreamCreateWithFlags(ss[k], cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->start, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->done, cudaEventDisableTiming);
  for (int k = 0; k < kPipeMaxChunks && e == cudaSuccess; ++k) {
    e = cudaEventCreateWithFlags(&p->ev_in[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_k[k], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&p->h_one, sizeof(int), cudaHostAllocDefault);
  if (e == cudaSuccess) *p->h_one = 1;
  if (e != cudaSuccess) { d3b200_pipe_destroy(p); return cuda_fail(e, "d3b200_pipe_create"); }
  *out = p;
  return D3B200_OK;
}
int d3b200_pipe_destroy(d3b200_pipe* p) {
  if (!p) return D3B200_OK;
  cudaStream_t ss[4] = {p->in, p->out, p->comp[0], p->comp[1]};
  for (int k = 0; k < 4; ++k) if (ss[k]) cudaStreamDestroy(ss[k]);
  if (p->start) cudaEventDestroy(p->start);
  if (p->done) cudaEventDestroy(p->done);
  for (int k = 0; k < kPipeMaxChunks; ++k) {
    if (p->ev_in[k]) cudaEventDestroy(p->ev_in[k]);
    if (p->ev_k[k]) cudaEventDestroy(p->ev_k[k]);
  }
  if (p->h_one) cudaFreeHost(p->h_one);
  free(p);
  return D3B200_OK;
}
int d3b200_batch_host_chunks(int nbatch, int chunk, int* bounds) {
  if (nbatch <= 0) return 0;
  int local[kPipeMaxChunks + 1];
  const int n = pipe_schedule(nbatch, chunk, local);
  if (bounds) for (int k = 0; k <= n; ++k) bounds[k] = local[k];
  return n;
}
size_t d3b200_batch_host_workspace_bytes_f64(int nbatch, int natoms, int with_g, int atm) {
  return host_workspace_bytes<double>(nbatch, natoms, with_g, atm);
}
size_t d3b200_batch_host_workspace_bytes_f32(int nbatch, int natoms, int with_g, int atm) {
  return host_workspace_bytes<float>(nbatch, natoms, with_g, atm);
}
int d3b200_batch_vjp_host_f64(d3b200_pipe* pipe, int nbatch, int natoms, const int64_t* h_numbers,
                              const double* h_positions, const d3b200_model_f64* model,
                              const d3b200_params* par, const double* h_g, double* h_energy,
                              double* h_grad, void* workspace, size_t workspace_bytes, int chunk,
                              void* stream) {
  return vjp_host_impl<double>(pipe, nbatch, natoms, h_numbers, h_positions, model, par, h_g, h_energy,
                               h_grad, workspace, workspace_bytes, chunk, stream);
}
int d3b200_batch_vjp_host_f32(d3b200_pipe* pipe, int nbatch, int natoms, const int64_t* h_numbers,
                              const float* h_positions, const d3b200_model_f32* model,
                              const d3b200_params* par, const float* h_g, float* h_energy,
                              float* h_grad, void* workspace, size_t workspace_bytes, int chunk,
                              void* stream) {
  return vjp_host_impl<float>(pipe, nbatch, natoms, h_numbers, h_positions, model, par, h_g, h_energy,
                              h_grad, workspace, workspace_bytes, chunk, stream);
}

}  // extern "C"
