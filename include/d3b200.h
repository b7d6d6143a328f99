/*
 * d3b200 -- C ABI of the B200-native DFT-D3 hot path (libd3b200.so).
 *
 * This is the drop-in boundary for tad-dftd3's pairwise / three-body path.
 * The reference has no FFI of its own (it is pure PyTorch); the boundary it
 * offers is the Python call surface
 *
 *     tad_dftd3.dftd3(numbers, positions, param, *, ref, rcov, rvdw, r4r2,
 *                     cutoff, ...)                      src/tad_dftd3/disp.py:79-165
 *
 * whose body (disp.py:150-165: cn_d3 -> weight_references -> atomic_c6 ->
 * dispersion) is what the entry points below replace.  The host side that
 * calls them is `tad_dftd3_b200/ops.py` (ctypes + torch.autograd.Function,
 * mirroring the reference's only custom op, model/c6.py:401-480).
 *
 * Conventions
 *  - all pointers are DEVICE pointers unless said otherwise; the library never
 *    allocates, never synchronises, keeps no mutable global state and launches
 *    on the stream it is given (a `cudaStream_t` passed as `void*`);
 *  - a padded batch is `numbers[nbatch][natoms]` (int64, 0 = padding or ghost
 *    atom) and `positions[nbatch][natoms][3]`, row-major, as produced by
 *    tad-mctc's `pack` and consumed by the reference; lengths in Bohr,
 *    energies in Hartree;
 *  - return value 0 = success, otherwise an error code; the message is
 *    available (per host thread) from `d3b200_last_error()`;
 *  - the `_f32` entry points take `float` arrays; scalar parameters are
 *    always passed as host doubles.
 */
#ifndef D3B200_H
#define D3B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D3B200_ABI_VERSION 2
#define D3B200_MAX_ELEMENT 104 /* reference defaults.py:72 */
#define D3B200_NREF 7          /* reference systems per element, reference.py:32-161 */

enum d3b200_status {
  D3B200_OK = 0,
  D3B200_EINVAL = 1,  /* bad argument (null pointer, size) */
  D3B200_ESIZE = 2,   /* structure too large for this entry point */
  D3B200_ECUDA = 3    /* CUDA runtime error, see d3b200_last_error() */
};

/* Damping parameters and cutoffs (host memory).
 * s6,s8,a1,a2: rational damping, reference damping/rational.py:38-68 and
 * disp.py:300-301; s9, rs9, alp: ATM term, disp.py:343-347, damping/atm.py;
 * cutoff: dispersion cutoff (disp.py:137, default 50); cn_cutoff, kcn: the
 * coordination number's own cutoff and steepness (tad-mctc defaults 25, 16). */
typedef struct d3b200_params {
  double s6, s8, a1, a2;
  double s9, rs9, alp;
  double cutoff, cn_cutoff, kcn;
  /* Model functions behind the reference's callable hooks (disp.py:89-91); 0 = the defaults.
   *  counting: 0 exp_count  1 / (1 + exp(-kcn (rc/r - 1)))            (tad-mctc ncoord/count.py)
   *            1 erf_count  (1 + erf(-kcn (r/rc - 1))) / 2
   *  damping:  0 rational_damping  1 / (r^n + (a1 sqrt(qq) + a2)^n)      (damping/rational.py:68)
   *            1 zero_damping      1 / (r^n (1 + 6 (rs_n sqrt(qq) / r)^alp_n)), rs_6 = a1, rs_8 = a2 of this
   *              struct, alp_6 = alp, alp_8 = alp + 2 (D3(0)-style damping on the radius sqrt(C8/C6))
   * Non-default functions run in the general fused kernel (structures up to d3b200_batch_max_atoms_*(1)). */
  int32_t counting, damping;
} d3b200_params;

/* Model data and per-atom inputs shared by all batch entry points.
 *  rcov, r4r2 : per atom `[nbatch][natoms]` (the reference's `rcov=`, `r4r2=`
 *               keyword arrays, disp.py:141-148) or, with *_per_element != 0,
 *               per element tables indexed by Z (length >= 104);
 *  refcn      : `[104][7]`, -1 = reference absent      (reference.py:32-161)
 *  refc6      : `[104][104][7][7]`                     (reference.py:164-187)
 *               must be symmetric, K[z1][z2][a][b] == K[z2][z1][b][a];
 *  rvdw       : ATM only. rvdw_mode 0: absent, 1: `[104][104]` element-pair
 *               table, 2: dense `[nbatch][natoms][natoms]` (the reference's
 *               `rvdw=` keyword, disp.py:143-146).
 *  batch_order: optional scheduling hint `[nbatch]` (a permutation of 0..nbatch-1,
 *               device or mapped host memory; NULL = identity): CTA k works on
 *               structure batch_order[k].  Largest structures first keeps the last
 *               wave of CTAs short; results do not depend on it. */
typedef struct d3b200_model_f64 {
  const double* rcov;
  const double* r4r2;
  int rcov_per_element;
  int r4r2_per_element;
  const double* refcn;
  const double* refc6;
  const double* rvdw;
  int rvdw_mode;
  const int32_t* batch_order;
} d3b200_model_f64;

typedef struct d3b200_model_f32 {
  const float* rcov;
  const float* r4r2;
  int rcov_per_element;
  int r4r2_per_element;
  const float* refcn;
  const float* refc6;
  const float* rvdw;
  int rvdw_mode;
  const int32_t* batch_order;
} d3b200_model_f32;

int d3b200_abi_version(void);
const char* d3b200_last_error(void);

/* Largest `natoms` the fused one-CTA-per-structure kernels accept
 * (order: 0 = energy / vjp, 1 = second order). */
int d3b200_batch_max_atoms_f64(int order);
int d3b200_batch_max_atoms_f32(int order);

/* Scratch the batch entry points need (device memory, caller-allocated, no
 * initialisation required): only the ATM term uses it, for the per-structure
 * sqrt|C6_ij| matrix.  order: 0 = energy / vjp, 1 = hvp; atm: par->s9 != 0. */
size_t d3b200_batch_workspace_bytes_f64(int nbatch, int natoms, int order, int atm);
size_t d3b200_batch_workspace_bytes_f32(int nbatch, int natoms, int order, int atm);

/* Atom-resolved D3(BJ)[+ATM] energies of a padded batch.
 * Replaces the body of `dftd3()`, reference disp.py:150-165 (+ disp.py:168-242).
 *   energy : `[nbatch][natoms]`, exact zeros on padding.
 * The ATM term is evaluated iff par->s9 != 0 (disp.py:234). */
int d3b200_batch_energy_f64(int nbatch, int natoms, const int64_t* numbers,
                            const double* positions, const d3b200_model_f64* model,
                            const d3b200_params* par, double* energy, void* workspace,
                            size_t workspace_bytes, void* stream);
int d3b200_batch_energy_f32(int nbatch, int natoms, const int64_t* numbers,
                            const float* positions, const d3b200_model_f32* model,
                            const d3b200_params* par, float* energy, void* workspace,
                            size_t workspace_bytes, void* stream);

/* Vector-Jacobian product of the above: with L = sum_i g[i] * E[i] (g == NULL: g = 1),
 *   grad  : dL/dpositions `[nbatch][natoms][3]`
 *   pgrad : dL/d(s6, s8, a1, a2, s9) per structure `[nbatch][5]`, or NULL.
 *   energy: optional `[nbatch][natoms]` (NULL to skip).
 * Replaces torch autograd through disp.py:150-165 (the analytic backward that
 * the reference only has for atomic_c6, model/c6.py:294-374). */
int d3b200_batch_vjp_f64(int nbatch, int natoms, const int64_t* numbers,
                         const double* positions, const d3b200_model_f64* model,
                         const d3b200_params* par, const double* g, double* energy,
                         double* grad, double* pgrad, void* workspace,
                         size_t workspace_bytes, void* stream);
int d3b200_batch_vjp_f32(int nbatch, int natoms, const int64_t* numbers,
                         const float* positions, const d3b200_model_f32* model,
                         const d3b200_params* par, const float* g, float* energy,
                         float* grad, float* pgrad, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Second order: directional derivative of (E, grad, pgrad) along the tangent
 * (dpositions, dpar->{s6,s8,a1,a2,s9}).  Because grad = dL/dx, the tangent of
 * grad is the Hessian-vector product of L; the tangent of E is the JVP of the
 * energies.  This is the double-backward that `examples/hessian.py:84-94`
 * (jacrev o jacrev) and `gradgradcheck` need.  dpar may be NULL (zero tangent);
 * any output may be NULL. */
int d3b200_batch_hvp_f64(int nbatch, int natoms, const int64_t* numbers,
                         const double* positions, const double* dpositions,
                         const d3b200_model_f64* model, const d3b200_params* par,
                         const d3b200_params* dpar, const double* g, double* grad,
                         double* pgrad, double* denergy, double* dgrad,
                         double* dpgrad, void* workspace, size_t workspace_bytes,
                         void* stream);
int d3b200_batch_hvp_f32(int nbatch, int natoms, const int64_t* numbers,
                         const float* positions, const float* dpositions,
                         const d3b200_model_f32* model, const d3b200_params* par,
                         const d3b200_params* dpar, const float* g, float* grad,
                         float* pgrad, float* denergy, float* dgrad, float* dpgrad,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * Host-buffer entry (the end-to-end form of `_batch_vjp_`): inputs and outputs
 * live in HOST memory (page-locked for the copies to be asynchronous), as a
 * caller of the reference's CPU `dftd3()` + autograd holds them.
 *
 * Three forms.  (1) Zero-copy, taken when chunk == 0, par->s9 == 0 and every host
 * buffer is page-locked and mapped (cudaHostAlloc / cudaHostRegister under UVA):
 * ONE launch of the fused kernel on `stream` reads the inputs and writes the
 * results directly in host memory; all its global accesses are coalesced, and
 * the PCIe traffic of each resident CTA overlaps the arithmetic of the others.
 * No workspace is used.  (2) Staged, otherwise (pageable memory, ATM, chunk > 0):
 * the batch is independent per structure, so it is cut into chunks of
 * `chunk` structures and pipelined over the streams of a `d3b200_pipe`:
 * H2D of chunk k+1, the fused kernel of chunk k (two alternating compute
 * streams, so that the tail of one launch overlaps the head of the next) and
 * D2H of chunk k-1 run concurrently.  Work is ordered after everything already
 * queued on `stream`, and `stream` waits for the last D2H: synchronise `stream`
 * before reading h_energy / h_grad.  Nothing is synchronised inside.
 * (3) Hybrid, chunk < 0 with page-locked mapped OUTPUT buffers (two-body + CN): the inputs are staged chunk by
 * chunk (|chunk| structures) with the copy engine while the kernels of the earlier chunks run and write their
 * results straight into host memory (posted PCIe writes of coalesced blocks): no device->host copy exists.
 *
 * A pipe owns 4 CUDA streams and a set of events on the current device; it is
 * the only object of this ABI that is created and destroyed, and one pipe must
 * not be used by two host threads at the same time.
 *
 *   model     : tables in DEVICE memory; per-element rcov / r4r2 only
 *               (rcov_per_element = r4r2_per_element = 1), rvdw_mode 0 or 1
 *   h_g       : cotangent `[nbatch][natoms]` or NULL (unit cotangent)
 *   h_energy  : `[nbatch][natoms]` or NULL;  h_grad: `[nbatch][natoms][3]`
 *   workspace : DEVICE staging, `d3b200_batch_host_workspace_bytes_*` bytes
 *   chunk     : > 0: staged form with this many structures per chunk;
 *               == 0: zero-copy when possible, else staged with about one wave of CTAs per chunk;
 *               < 0: hybrid form with |chunk| structures per chunk when possible, else staged */
typedef struct d3b200_pipe d3b200_pipe;
int d3b200_pipe_create(d3b200_pipe** pipe);
int d3b200_pipe_destroy(d3b200_pipe* pipe);
/* Chunk schedule of the staged form (= its number of kernel launches): returns the number of chunks
 * and, if `bounds` is not NULL, writes the chunk boundaries bounds[0..n] (room for 65 ints). */
int d3b200_batch_host_chunks(int nbatch, int chunk, int* bounds);
size_t d3b200_batch_host_workspace_bytes_f64(int nbatch, int natoms, int with_g, int atm);
size_t d3b200_batch_host_workspace_bytes_f32(int nbatch, int natoms, int with_g, int atm);
int d3b200_batch_vjp_host_f64(d3b200_pipe* pipe, int nbatch, int natoms, const int64_t* h_numbers,
                              const double* h_positions, const d3b200_model_f64* model,
                              const d3b200_params* par, const double* h_g, double* h_energy,
                              double* h_grad, void* workspace, size_t workspace_bytes, int chunk,
                              void* stream);
int d3b200_batch_vjp_host_f32(d3b200_pipe* pipe, int nbatch, int natoms, const int64_t* h_numbers,
                              const float* h_positions, const d3b200_model_f32* model,
                              const d3b200_params* par, const float* h_g, float* h_energy,
                              float* h_grad, void* workspace, size_t workspace_bytes, int chunk,
                              void* stream);

/* ------------------------------------------------------------------------
 * One large structure, tiled over the whole GPU and shardable by rows.
 *
 * Same path as above (reference disp.py:150-165) for structures too large for
 * one CTA (BASELINE config 4: 50 k atoms) and for the i-row sharded multi-GPU
 * mode.  The caller (tad_dftd3_b200/system.py) removes ghost atoms, sorts the
 * atoms spatially and by element inside each 128-atom tile, pads to a multiple
 * of 128 with far-away Z = 0 atoms and supplies:
 *   atoms [natoms][4]  x, y, z, kcn * rcov          aux [natoms][2]  sqrt(r4r2), g
 *   z     [natoms]     int32 atomic numbers          box [natoms/32][6]  min xyz, max xyz
 * Sweeps (each writes only rows row_begin..row_end-1, multiples of 128; all
 * other arrays are read for every atom, so on several GPUs `cn` and `decn`
 * must be all-gathered between the sweeps):
 *   _cn       cn[i]                              tad-mctc cn_d3, call site disp.py:150-152
 *   _weights  w[i][8], dw[i][8] for ALL atoms    model/weights.py:104-176
 *   _pair     energy[i], grad[i][3] (direct part), decn[i] = dL/dCN_i
 *             (want_grad = 0: energy only)        model/c6.py + disp.py:276-302
 *   _cnbwd    grad[i][3] += CN chain-rule term    (autograd through cn_d3)
 * g (in aux) is the cotangent of the per-atom energies, L = sum g_i E_i.
 *   _tvec     (optional, after _weights; needs nelem/zlist/eidx/tvec) pre-contracts
 *             T_k[e][a] = sum_b K[Z_e,Z_k,a,b] w_kb (and Td_k with dw_k) for every atom; `_pair` and `_atm`
 *             then need one 7-wide dot per C6 instead of a 7x7 contraction per
 *             (row, partner element).
 * symmetric = 1 (needs the T vectors for `_pair`): every unordered pair is evaluated
 *             once; a launch owns the row tiles row_begin/128 + k*tile_stride below
 *             row_end and walks the tiles J >= I; results are ADDED with fp64
 *             atomics to cn / energy / grad / decn for BOTH atoms of a pair, so the
 *             arrays must be zero before `_cn` (cn) and before `_pair` (the others), and
 *             on several GPUs they are summed (all-reduce) instead of gathered.
 * jsplit >= 1 CTAs share each 128-row tile (the j-tiles are dealt round-robin
 * among them, which is what fills the GPU when N / 128 is small compared with
 * the SM count); their partial rows go through `scratch`
 * (5 * jsplit * natoms reals) and are summed in a fixed order. */
typedef struct d3b200_system_f64 {
  int natoms, row_begin, row_end;
  const double* atoms;
  const double* aux;
  const int32_t* z;
  const double* box;
  const double* refcn;
  const double* refc6;
  double* cn;
  double* w;
  double* dw;
  double* decn;
  double* energy;
  double* grad;
  int jsplit;
  double* scratch;
  int nelem;            /* optional fast path, 0 = off: number of distinct elements (<= 8) */
  const int32_t* zlist; /* [nelem] the distinct atomic numbers                              */
  const int8_t* eidx;   /* [natoms] index of each atom's element in zlist, -1 = padding     */
  double* tvec;            /* [natoms*nelem*16 + 64] scratch filled by _system_tvec_           */
  int symmetric;        /* 1: pairs-once sweeps (see below); 0: full-row sweeps             */
  int tile_stride;      /* symmetric only: own row tiles row_begin/128 + k*tile_stride      */
  void* atm_work;       /* optional scratch of `_system_atm_` (d3b200_system_atm_workspace_bytes_*) */
  size_t atm_work_bytes;
} d3b200_system_f64;

typedef struct d3b200_system_f32 {
  int natoms, row_begin, row_end;
  const float* atoms;
  const float* aux;
  const int32_t* z;
  const float* box;
  const float* refcn;
  const float* refc6;
  float* cn;
  float* w;
  float* dw;
  float* decn;
  float* energy;
  float* grad;
  int jsplit;
  float* scratch;
  int nelem;            /* optional fast path, 0 = off: number of distinct elements (<= 8) */
  const int32_t* zlist; /* [nelem] the distinct atomic numbers                              */
  const int8_t* eidx;   /* [natoms] index of each atom's element in zlist, -1 = padding     */
  float* tvec;            /* [natoms*nelem*16 + 64] scratch filled by _system_tvec_           */
  int symmetric;        /* 1: pairs-once sweeps (see below); 0: full-row sweeps             */
  int tile_stride;      /* symmetric only: own row tiles row_begin/128 + k*tile_stride      */
  void* atm_work;       /* optional scratch of `_system_atm_` (d3b200_system_atm_workspace_bytes_*) */
  size_t atm_work_bytes;
} d3b200_system_f32;

int d3b200_system_cn_f64(const d3b200_system_f64* sys, const d3b200_params* par, void* stream);
int d3b200_system_weights_f64(const d3b200_system_f64* sys, const d3b200_params* par, void* stream);
int d3b200_system_tvec_f64(const d3b200_system_f64* sys, const d3b200_params* par, void* stream);
int d3b200_system_pair_f64(const d3b200_system_f64* sys, const d3b200_params* par, int want_grad, void* stream);
int d3b200_system_cnbwd_f64(const d3b200_system_f64* sys, const d3b200_params* par, void* stream);
/* Packs the position-dependent arrays of a plan in one launch (instead of a dozen tensor
 * operations on the host side): atoms[natoms][4] = {positions[perm[i]] for i < nreal, else the
 * far-away padding position of the template; kcn*rcov from the template} and
 * box[natoms/32][6] = bounding box of every 32-atom sub-tile over its real atoms (z != 0; a
 * sub-tile of padding only: its first atom).  perm: sorted slot -> original atom index. */
int d3b200_system_pack_f64(int natoms, int nreal, const int64_t* perm, const double* positions,
                           const double* atoms_template, const int32_t* z, double* atoms, double* box,
                           void* stream);
int d3b200_system_pack_f32(int natoms, int nreal, const int64_t* perm, const float* positions,
                           const float* atoms_template, const int32_t* z, float* atoms, float* box,
                           void* stream);

/* Damping-parameter gradient of the two-body term for the tiled path (needs the T vectors):
 * writes the per-row sums scratch[q][jsplit][natoms], q = 0..3 <-> s6, s8, a1, a2
 * (sum_j C6 dP/dq of rows row_begin..row_end-1; all other scratch entries untouched);
 * dL/dq = sum_i -1/2 g_i sum_k scratch[q][k][i].  Reference: autograd through
 * disp.py:300-301, damping/rational.py:68. */
int d3b200_system_pgrad_f64(const d3b200_system_f64* sys, const d3b200_params* par, void* stream);
int d3b200_system_pgrad_f32(const d3b200_system_f32* sys, const d3b200_params* par, void* stream);
int d3b200_system_cn_f32(const d3b200_system_f32* sys, const d3b200_params* par, void* stream);
int d3b200_system_weights_f32(const d3b200_system_f32* sys, const d3b200_params* par, void* stream);
int d3b200_system_tvec_f32(const d3b200_system_f32* sys, const d3b200_params* par, void* stream);
int d3b200_system_pair_f32(const d3b200_system_f32* sys, const d3b200_params* par, int want_grad, void* stream);
int d3b200_system_cnbwd_f32(const d3b200_system_f32* sys, const d3b200_params* par, void* stream);

/* ATM three-body term of the tiled path (reference damping/atm.py:86-155, ordered
 * cutoff rule kept).  Run after `_pair` and before `_cnbwd`: ADDS (fp64 atomics,
 * once per short edge) the three-body parts to energy[], grad[] and decn[] for
 * the triples whose owner edge (i, j), i < j, has i in row_begin..row_end-1
 * (any row range).  rvdw: `[104][104]` element-pair table.  Uses the T vectors
 * when the system struct carries them (after `_system_tvec_`).
 * With the T vectors AND `atm_work` (device scratch of at least
 * `d3b200_system_atm_workspace_bytes_*(natoms)` bytes, no initialisation needed)
 * every triple is evaluated once: persistent CTAs take centre atoms from a
 * counter, compact the centre's neighbour list into the scratch and walk the
 * pairs of neighbours (see csrc/d3_atm3.cuh).  The launch owns the centres of
 * the 32-atom sub-tiles row_begin/32 + k*tile_stride below row_end (row range
 * aligned to 32; tile_stride 0 or 1: the whole range), or, with tile_stride < 0, the
 * single atoms row_begin + k*|tile_stride| below row_end (any row_begin), which is how
 * ranks shard the term: dealt atom by atom the ranks' costs differ by well under 1 %.
 * `_workspace_bytes_` includes, while it stays below $D3B200_ATM_PAIRTAB_MAX_BYTES (default
 * 8 GiB; 0 = never), a dense pair table of 4 reals per atom pair (32 natoms^2 bytes in fp64)
 * that the launch fills itself: everything of a triple that depends on one edge alone is then
 * computed once per pair instead of once per triple.  A smaller `atm_work` (at least the
 * neighbour scratch) selects the table-free sweep; the results agree to rounding. */
size_t d3b200_system_atm_workspace_bytes_f64(int natoms);
size_t d3b200_system_atm_workspace_bytes_f32(int natoms);
int d3b200_system_atm_f64(const d3b200_system_f64* sys, const d3b200_params* par,
                          const double* rvdw, int want_grad, void* stream);
int d3b200_system_atm_f32(const d3b200_system_f32* sys, const d3b200_params* par,
                          const float* rvdw, int want_grad, void* stream);

/* ------------------------------------------------------------------------
 * Stage-level entry points: the reference's operator API keeps
 * the intermediates as dense tensors (README.rst:238-261, test/test_disp/
 * test_disp.py feeds an external c6).  Padded-batch layout as above.
 *   _stage_cn      cn[B][N]          tad-mctc cn_d3 / exp_count (rcov per atom [B][N])
 *   _stage_weights w[n][7]           model/weights.py:104-176 (n = B*N atoms, flat)
 *   _stage_c6      c6[B][N][N]       model/c6.py:186-217
 *   _stage_disp2   energy[B][N]      disp.py:276-302 with dense c6 and per-atom r4r2
 *   _stage_atm     energy[B][N]      damping/atm.py:86-155 with dense c6 and rvdw [B][N][N] */
int d3b200_stage_cn_f64(int nbatch, int natoms, const int64_t* numbers, const double* positions,
                        const double* rcov, double cutoff, double kcn, double* cn, void* stream);
int d3b200_stage_weights_f64(size_t n, const int64_t* numbers, const double* cn, const double* refcn,
                             double* weights, void* stream);
int d3b200_stage_c6_f64(int nbatch, int natoms, const int64_t* numbers, const double* weights,
                        const double* refc6, double* c6, void* stream);
/* Derivatives of `_stage_c6_` (the reference's hand-written AtomicC6 backward,
 * model/c6.py:294-374).  C6 is bilinear in the weights, so two primitives close the
 * chain to any order (tad_dftd3_b200/stages.py wires them into torch.autograd):
 *   _stage_c6_bilinear  out[B][N][N]: out_ij = sum_ab u_ia K[Zi,Zj,a,b] w_jb
 *   _stage_c6_contract  out[B][N][7]: out_ia = sum_j G_ij sum_b K[Zi,Zj,a,b] w_jb */
int d3b200_stage_c6_bilinear_f64(int nbatch, int natoms, const int64_t* numbers, const double* u,
                                 const double* w, const double* refc6, double* out, void* stream);
int d3b200_stage_c6_contract_f64(int nbatch, int natoms, const int64_t* numbers, const double* G,
                                 const double* w, const double* refc6, double* out, void* stream);
int d3b200_stage_c6_bilinear_f32(int nbatch, int natoms, const int64_t* numbers, const float* u,
                                 const float* w, const float* refc6, float* out, void* stream);
int d3b200_stage_c6_contract_f32(int nbatch, int natoms, const int64_t* numbers, const float* G,
                                 const float* w, const float* refc6, float* out, void* stream);
int d3b200_stage_disp2_f64(int nbatch, int natoms, const int64_t* numbers, const double* positions,
                           const double* c6, const double* r4r2, const d3b200_params* par,
                           double* energy, void* stream);
int d3b200_stage_atm_f64(int nbatch, int natoms, const int64_t* numbers, const double* positions,
                         const double* c6, const double* rvdw, const d3b200_params* par,
                         double* energy, void* stream);
/* Vector-Jacobian products of the stage functions (what torch autograd derives for the reference's stage
 * pipeline, README.rst:238-261 under `autograd.grad`).  Cotangents in, gradients out, first order:
 *   _stage_cn_bwd       gx[B][N][3]   from gcn[B][N]              (positions; rcov is a constant)
 *   _stage_weights_bwd  gcn[n]        from gw[n][7]
 *   _stage_disp2_bwd    gx[B][N][3], gc6[B][N][N] (entries of the dense c6 are independent variables, as in the
 *                       reference), gpar[B][N][4] = ge_i dE_i/d(s6, s8, a1, a2) (sum over atoms = parameter gradient)
 *                       from ge[B][N]
 *   _stage_atm_bwd      gx[B][N][3], gc6[B][N][N] from ge[B][N]; both outputs must be ZERO on entry (fp64 atomics);
 *                       the ordered term (i; j, k) uses c6[i][j], c6[i][k], c6[j][k] as damping/atm.py:93-97 indexes them */
/* `_stage_cn_` / `_stage_cn_bwd_` with the counting function chosen by id (d3b200_params.counting: 0 exp_count,
 * 1 erf_count); `_stage_disp2_` / `_stage_disp2_bwd_` read the damping function from par->damping (and par->alp). */
int d3b200_stage_cn_model_f64(int nbatch, int natoms, const int64_t* numbers, const double* positions,
                              const double* rcov, double cutoff, double kcn, int counting, double* cn, void* stream);
int d3b200_stage_cn_model_bwd_f64(int nbatch, int natoms, const int64_t* numbers, const double* positions,
                                  const double* rcov, double cutoff, double kcn, int counting, const double* gcn,
                                  double* gx, void* stream);
int d3b200_stage_cn_model_f32(int nbatch, int natoms, const int64_t* numbers, const float* positions,
                              const float* rcov, double cutoff, double kcn, int counting, float* cn, void* stream);
int d3b200_stage_cn_model_bwd_f32(int nbatch, int natoms, const int64_t* numbers, const float* positions,
                                  const float* rcov, double cutoff, double kcn, int counting, const float* gcn,
                                  float* gx, void* stream);
int d3b200_stage_cn_bwd_f64(int nbatch, int natoms, const int64_t* numbers, const double* positions,
                            const double* rcov, double cutoff, double kcn, const double* gcn, double* gx, void* stream);
int d3b200_stage_weights_bwd_f64(size_t n, const int64_t* numbers, const double* cn, const double* refcn,
                                 const double* gw, double* gcn, void* stream);
int d3b200_stage_disp2_bwd_f64(int nbatch, int natoms, const int64_t* numbers, const double* positions,
                               const double* c6, const double* r4r2, const d3b200_params* par, const double* ge,
                               double* gx, double* gc6, double* gpar, void* stream);
int d3b200_stage_atm_bwd_f64(int nbatch, int natoms, const int64_t* numbers, const double* positions,
                             const double* c6, const double* rvdw, const d3b200_params* par, const double* ge,
                             double* gx, double* gc6, void* stream);
int d3b200_stage_cn_bwd_f32(int nbatch, int natoms, const int64_t* numbers, const float* positions,
                            const float* rcov, double cutoff, double kcn, const float* gcn, float* gx, void* stream);
int d3b200_stage_weights_bwd_f32(size_t n, const int64_t* numbers, const float* cn, const float* refcn,
                                 const float* gw, float* gcn, void* stream);
int d3b200_stage_disp2_bwd_f32(int nbatch, int natoms, const int64_t* numbers, const float* positions,
                               const float* c6, const float* r4r2, const d3b200_params* par, const float* ge,
                               float* gx, float* gc6, float* gpar, void* stream);
int d3b200_stage_atm_bwd_f32(int nbatch, int natoms, const int64_t* numbers, const float* positions,
                             const float* c6, const float* rvdw, const d3b200_params* par, const float* ge,
                             float* gx, float* gc6, void* stream);
int d3b200_stage_cn_f32(int nbatch, int natoms, const int64_t* numbers, const float* positions,
                        const float* rcov, double cutoff, double kcn, float* cn, void* stream);
int d3b200_stage_weights_f32(size_t n, const int64_t* numbers, const float* cn, const float* refcn,
                             float* weights, void* stream);
int d3b200_stage_c6_f32(int nbatch, int natoms, const int64_t* numbers, const float* weights,
                        const float* refc6, float* c6, void* stream);
int d3b200_stage_disp2_f32(int nbatch, int natoms, const int64_t* numbers, const float* positions,
                           const float* c6, const float* r4r2, const d3b200_params* par,
                           float* energy, void* stream);
int d3b200_stage_atm_f32(int nbatch, int natoms, const int64_t* numbers, const float* positions,
                         const float* c6, const float* rvdw, const d3b200_params* par,
                         float* energy, void* stream);

/* Measurement utility (bench.py): launches `blocks` CTAs of 256 threads, each
 * thread running `iters` x 64 dependent-chain FMAs (8 independent chains), i.e.
 * blocks * 256 * iters * 128 flops.  `out` is one device scalar (never written
 * in practice).  Gives the sustained FP64 / FP32 FMA-pipe rate that the pair
 * kernels' roofline is quoted against. */
int d3b200_probe_fma_f64(int blocks, int iters, double* out, void* stream);
int d3b200_probe_fma_f32(int blocks, int iters, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* D3B200_H */
