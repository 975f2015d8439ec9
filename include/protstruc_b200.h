/*
 * protstruc_b200 — C-ABI of the B200 (sm_100a) geometric-feature hot path.
 *
 * This is the drop-in boundary: every entry point below replaces one method (or
 * free function) of the reference's Python surface, cited as file:line into the
 * reference tree (dohlee/protstruc).  The reference has no FFI of its own (it is
 * pure Python), so these are the functions a maintainer would bind with ctypes
 * (see INTEGRATION.md).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *   - every pointer is a DEVICE pointer into memory owned by the caller
 *     (the Python side allocates outputs with torch and passes data_ptr());
 *     kernels never allocate or free device memory;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *     launches are asynchronous on that stream, no hidden synchronisation;
 *   - every function returns PS_OK (0) or a negative ps_status code and records a
 *     human-readable message retrievable with ps_last_error_string() (thread local);
 *   - coordinates are fp32, row-major contiguous (B, L, A, 3);
 *   - masks: `mask_dtype` is PS_MASK_BOOL (1 byte per element, 0/1) or
 *     PS_MASK_F32 (fp32, arbitrary values, multiplied like the reference does).
 */
#ifndef PROTSTRUC_B200_H_
#define PROTSTRUC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ps_status {
    PS_OK = 0,
    PS_ERR_BAD_SHAPE = -1,     /* a dimension is <= 0 or exceeds a kernel limit      */
    PS_ERR_NULL_POINTER = -2,  /* a required pointer is NULL                         */
    PS_ERR_BAD_DTYPE = -3,     /* unknown mask_dtype / kind code                     */
    PS_ERR_BAD_SLOT = -4,      /* atom slot index outside [0, A)                     */
    PS_ERR_MISALIGNED = -5,    /* pointer alignment requirement violated             */
    PS_ERR_CUDA = -6           /* CUDA runtime error (message has cudaGetErrorString) */
} ps_status;

enum { PS_MASK_BOOL = 0, PS_MASK_F32 = 1 };
enum { PS_ANGLE_DIHEDRAL = 0, PS_ANGLE_PLANAR = 1 };

/* Library / build identification. */
int ps_abi_version(void);                 /* bumps on any signature change            */
const char* ps_build_info(void);          /* "sm_100a nvcc <ver> ..."                  */
const char* ps_last_error_string(void);   /* message of the last failing call (thread) */
int ps_device_sm_count(int device);       /* >0, or negative ps_status                 */
/*
 * Leaves `n` SMs of the device to other work: every launcher sizes its grid for (SM count - n) SMs from now on
 * (process-wide; 0 restores the default).  The fused tile kernels are persistent, one CTA per SM with most of its
 * shared memory and registers, so a collective on another stream (the optional NCCL gather of compact features) can
 * only run BESIDE them on SMs they leave free.  Returns the previous setting, or a negative ps_status.
 */
int ps_reserve_sms(int n);

/*
 * K1 — all-atom pairwise distance matrix with the pair mask fused in.
 * Replaces StructureBatch.pairwise_distance_matrix  (protstruc/protstruc.py:455-484):
 *   dist[b,i,j,a,c]      = || xyz[b,i,a,:] - xyz[b,j,c,:] ||_2           (NOT masked; NaN flows)
 *   dist_mask[b,i,j,a,c] = atom_mask[b,i,a] * atom_mask[b,j,c]            (dtype of atom_mask)
 * xyz (B,L,A,3) f32; atom_mask (B,L,A) of mask_dtype; dist (B,L,L,A,A) f32;
 * dist_mask (B,L,L,A,A) of mask_dtype.  atom_mask/dist_mask may both be NULL
 * (distances only).  dist may be NULL (mask only).  Any A, L and (naturally
 * aligned) output pointers are accepted: A in {5, 10, 14, 15} with 16-byte
 * aligned outputs takes the staged kernel, every other case the any-A tile
 * kernel (plain stores instead of TMA bulk stores where alignment forbids them).
 */
int ps_pair_dist_mask(const float* xyz, const void* atom_mask, int mask_dtype,
                      float* dist, void* dist_mask,
                      int B, int L, int A, void* stream);

/*
 * K2 — pairwise dihedral / planar angle between residues for arbitrary atom-slot lists.
 * Replaces StructureBatch.pairwise_dihedrals / pairwise_planar_angles
 * (protstruc/protstruc.py:620-660) including the gather of _pairwise_xyz (:589-618)
 * and geometry.dihedral / geometry.angle (protstruc/geometry.py:39-124).
 *   kind = PS_ANGLE_DIHEDRAL: n_i + n_j == 4;  kind = PS_ANGLE_PLANAR: n_i + n_j == 3
 *   the point list is slots_i of residue i followed by slots_j of residue j
 * out (B,L,L) f32.
 */
int ps_pair_angles(const float* xyz, int B, int L, int A,
                   const int* slots_i, int n_i, const int* slots_j, int n_j,
                   int kind, float* out, void* stream);
/* Comparison hook: variant 0 = default (packed-FP32 kernel when the points come from both residues: NaN placement of
 * the reference, <= 1e-5 rad where min sin(bond angle) >= 0.1), 1 = the exact-operation-sequence kernel. */
int ps_pair_angles_ex(const float* xyz, int B, int L, int A,
                      const int* slots_i, int n_i, const int* slots_j, int n_j,
                      int kind, float* out, int variant, void* stream);

/*
 * K2f — the trRosetta-style triple in one pass over the pairs.
 * Replaces the three angle calls of StructureBatch.inter_residue_geometry
 * (protstruc/protstruc.py:810-815), exactly as the reference defines them:
 *   omega[b,i,j] = dihedral(CA_i, CB_i, CA_j, CB_j)
 *   theta[b,i,j] = dihedral(N_i,  CA_i, CB_i, CB_j)
 *   phi  [b,i,j] = angle   (CA_i, CB_i, CB_j)
 * virtual_cb != 0: CB is recomputed in-register from N, CA, C with the
 * ideal-geometry coefficients of protstruc/geometry.py:217-221 instead of read from slot 4.
 * Any of omega/theta/phi may be NULL.
 */
int ps_trrosetta_angles(const float* xyz, int B, int L, int A, int virtual_cb,
                        float* omega, float* theta, float* phi, void* stream);
/* Tuning / comparison hook: variant 0 = default (packed-FP32 kernel compiled for 3 CTAs per SM), 1 = the
 * exact-operation-sequence kernel, 4 / 5 / 6 = the packed kernel compiled for 4 / 5 / 6 CTAs per SM, 3 = two rows
 * per loop iteration at 2 CTAs per SM.  All packed variants produce the same bits. */
int ps_trrosetta_angles_ex(const float* xyz, int B, int L, int A, int virtual_cb,
                           float* omega, float* theta, float* phi, int variant, void* stream);

/*
 * K1+K2f — the full pairwise feature set in ONE kernel: distance matrix, pair mask
 * and omega/theta/phi.  Replaces StructureBatch.inter_residue_geometry
 * (protstruc/protstruc.py:790-817); d_ca/d_cb/d_no are views of `dist` taken by the caller.
 * Requires A >= 5.  One launch when A is a staged atom count (5, 10, 14, 15), L is at least the
 * pairs per tile (128, 64, 32, 32), the outputs are 16-byte aligned and mask_dtype is PS_MASK_BOOL;
 * otherwise the distance / mask / angle kernels run back to back on `stream` with identical results.
 */
int ps_inter_residue_geometry(const float* xyz, const void* atom_mask, int mask_dtype,
                              float* dist, void* dist_mask,
                              float* omega, float* theta, float* phi,
                              int B, int L, int A, void* stream);

/*
 * The same launch with the COMPACT features gathered on the way: compact is a contiguous (6, B, L, L) f32 buffer that
 * receives omega, theta, phi and d_ca = dist[..,CA,CA], d_cb = dist[..,CB,CB], d_no = dist[..,N,O] — the (B, L, L)
 * planes of inter_residue_geometry (protstruc/protstruc.py:801-815) as dense tensors, which is what the optional
 * multi-GPU exchange all-gathers over NVLink (the reference returns the distance planes as strided views of `dist`;
 * the fused kernel reads them back from the finished tile in shared memory, 12 extra bytes per residue pair).
 */
int ps_inter_residue_geometry_compact(const float* xyz, const void* atom_mask, int mask_dtype,
                                      float* dist, void* dist_mask, float* compact,
                                      int B, int L, int A, void* stream);

/*
 * The same launch FUSED WITH THE ALL-GATHER of the compact features over NVLink / NVSwitch peer memory (optional
 * exchange step of a batch-sharded job; one process per GPU).  Every rank owns a symmetric buffer
 * gathered (6, world, shard, L, L) f32 — feature-major: omega, theta, phi, d_ca, d_cb, d_no; then rank; then the
 * rank's `shard` structures — and passes the addresses of ALL ranks' buffers as mapped into its own address space
 * (peer_buffers[0 .. world), its own included; e.g. torch.distributed._symmetric_memory's buffer_ptrs), or
 * additionally ONE NVSwitch multicast address of the buffer (multicast_buffer, may be NULL).  The angle warp of the tile
 * kernel stores the six values of its residue pair straight into rank `rank`'s slab of every peer buffer — one
 * multimem.st through the switch when a multicast address is given, else one store per peer — so the transfer rides
 * on the kernel tile by tile; there is no separate collective and no staging buffer.  The caller makes the stores
 * visible with a barrier across the ranks after the kernel (symmetric memory's barrier, or any collective).
 * B <= shard structures of this rank; needs the linear-sweep kernel (A = 15, L >= 32, 16-byte aligned arrays),
 * PS_ERR_BAD_SHAPE otherwise (use ps_inter_residue_geometry_compact + an NCCL all-gather there).
 */
int ps_inter_residue_geometry_push(const float* xyz, const void* atom_mask, int mask_dtype,
                                   float* dist, void* dist_mask,
                                   void* const* peer_buffers, int world, int rank, void* multicast_buffer,
                                   int shard, int B, int L, int A, void* stream);

/*
 * K3 — per-residue backbone features.
 * Replaces StructureBatch.backbone_dihedrals (protstruc/protstruc.py:486-541) with the
 * terminal masks of :435-453, and StructureBatch.backbone_orientations (:543-571) →
 * geometry.gram_schmidt (protstruc/geometry.py:413-439).
 *   residue_mask (B,L) uint8 0/1;  chain_idx (B,L) f32, NaN = padding
 *   dihedrals (B,L,3) f32 [phi,psi,omega]; dihedral_mask (B,L,3) uint8
 *   frames (B,L,3,3) f32, columns e1,e2,e3 from slots (a1,a2,a3)
 * Either the (dihedrals, dihedral_mask) pair or frames may be NULL.
 * residue_mask / chain_idx are only required when dihedrals are requested.
 */
int ps_backbone(const float* xyz, const uint8_t* residue_mask, const float* chain_idx,
                int B, int L, int A, int a1, int a2, int a3,
                float* dihedrals, uint8_t* dihedral_mask, float* frames, void* stream);

/*
 * K4 — masked per-structure, per-axis mean / population std, optionally applied.
 * Replaces StructureBatch.standardize (protstruc/protstruc.py:696-734), with the
 * per-structure broadcast the reference intends.
 *   mu, sd (B,3) f32 out;  xyz_out (B,L,A,3) f32 = (xyz - mu) / sd, may be NULL; must NOT alias xyz (the kernels
 *   read xyz through the read-only path).
 */
int ps_masked_stats(const float* xyz, const void* atom_mask, int mask_dtype,
                    int B, int L, int A, float* mu, float* sd, float* xyz_out, void* stream);
/* Comparison hook: variant 0 = default (register-resident single-read kernels), 1 = the three-pass kernel of round 1,
 * 2 = the scalar-mapped register-resident kernel also where the quad kernel applies, 3 = the quad kernel restricted to
 * one or two 4-atom groups per thread (without the dense configuration that keeps every structure resident). */
int ps_masked_stats_ex(const float* xyz, const void* atom_mask, int mask_dtype,
                       int B, int L, int A, float* mu, float* sd, float* xyz_out, int variant, void* stream);

/*
 * Affine per-structure map used by unstandardize (protstruc/protstruc.py:736-744):
 *   out = xyz * scale[b,:] + shift[b,:]   (two separately rounded fp32 ops)
 */
int ps_scale_shift(const float* xyz, const float* scale, const float* shift,
                   int B, int L, int A, float* xyz_out, void* stream);

/*
 * NaN-skipping mean of one atom slot over residues.
 * Replaces StructureBatch.center_of_mass (protstruc/protstruc.py:746-757).  out (B,3) f32.
 */
int ps_center_of_mass(const float* xyz, int B, int L, int A, int slot, float* out, void* stream);

/*
 * Per-structure translation, used by center_at / translate (protstruc/protstruc.py:662-679, 759-788):
 *   out[b,l,a,:] = xyz[b,l,a,:] + t[b,:]   (t has B rows, or 1 row when t_rows == 1)
 */
int ps_translate(const float* xyz, const float* t, int t_rows,
                 int B, int L, int A, float* xyz_out, void* stream);

/*
 * Rigid-frame family.
 *  ps_local_xyz — StructureBatch.get_local_xyz (protstruc/protstruc.py:347-362):
 *      out[b,l,a,:] = R_{b,l}^T xyz[b,l,a,:] - xyz[b,l,ca_slot,:], R from Gram-Schmidt on slots (a1,a2,a3)
 *  ps_rotate — StructureBatch.rotate (protstruc/protstruc.py:681-694): xyz_out = R_b xyz, rotation is
 *      (rot_rows,3,3) with rot_rows = B or 1; xyz_out must not alias xyz
 *  ps_frames_to_backbone — StructureBatch.from_backbone_orientations_translations
 *      (protstruc/protstruc.py:263-319): xyz[b,l,a,:] = R_{b,l} ideal[a,:] + t_{b,l} for a < n_ideal,
 *      0 otherwise; atom_mask (B,L,A) f32 = 1 for the placed atoms.  `ideal` is (n_ideal,3) f32 on the device.
 *  ps_translate_bcast — StructureBatch.translate (protstruc/protstruc.py:662-679): xyz_out = xyz + t where
 *      t is addressed as t[b*stride_b + l*stride_l + a*stride_a + k] (strides in elements, 0 = broadcast).
 */
int ps_local_xyz(const float* xyz, int B, int L, int A, int a1, int a2, int a3, int ca_slot,
                 float* out, void* stream);
int ps_rotate(const float* xyz, const float* rotation, int rot_rows, int B, int L, int A,
              float* xyz_out, void* stream);
int ps_frames_to_backbone(const float* orientations, const float* translations, const float* ideal,
                          int n_ideal, int B, int L, int A, float* xyz, float* atom_mask, void* stream);
int ps_translate_bcast(const float* xyz, const float* t, int64_t stride_b, int64_t stride_l,
                       int64_t stride_a, int B, int L, int A, float* xyz_out, void* stream);

/*
 * Batched Kabsch — the solve inside StructureBatch.align (protstruc/protstruc.py:880-918,
 * geometry.kabsch protstruc/geometry.py:442-480): for every structure the rotation / translation that best maps
 * the selected atoms of `source` onto `target`:  rotation (B,3,3), translation (B,3) with
 * target ~= rotation @ source + translation.  source (B,n_atoms,3), target (target_rows,n_atoms,3) with
 * target_rows = B or 1, mask (B,n_atoms) uint8.
 */
int ps_kabsch(const float* source, const float* target, const uint8_t* mask, int target_rows, int B,
              int n_atoms, float* rotation, float* translation, void* stream);

/*
 * Top-k nearest residues — StructureBatch.get_topk_nearest_residue_mask
 * (protstruc/protstruc.py:819-862), one structure: CA distance to the closest of n_query points, residues
 * with valid == 0 pushed to 1e9, the k smallest marked.  scratch: L floats of device workspace.  out (L) uint8.
 */
int ps_topk_nearest_residue_mask(const float* xyz, const uint8_t* valid, const float* query, int n_query,
                                 int L, int A, int ca_slot, int k, float* scratch, uint8_t* out,
                                 void* stream);

/*
 * PDB text ingest — HOST pointers, no GPU work.  Replaces PDB.read_pdb / tidy_structure /
 * PDB._initialize_lookup / PDB._compute_atom_xyz (protstruc/pdb.py:24-40, 55-151) without biotite.
 * Parses one PDB file held in memory into per-residue rows: xyz (capacity,15,3) f32 NaN-filled, atom_mask
 * (capacity,15) uint8, chain_idx / residue_number (capacity) int32, chain_id / insertion_code / one_letter
 * (capacity) chars (insertion 0 = none; one_letter 'X' = UNK gap placeholder).  *n_residues receives the row
 * count; call with capacity = 0 (arrays may be NULL) to size the buffers first.
 */
int ps_host_pdb_parse(const char* text, int64_t len, int capacity, float* xyz, uint8_t* atom_mask,
                      int32_t* chain_idx, char* chain_id, int32_t* residue_number, char* insertion_code,
                      char* one_letter, int* n_residues);

/*
 * Host-buffer pipeline: the reference-facing call with HOST arrays in and out (what a caller of
 * StructureBatch.inter_residue_geometry, protstruc/protstruc.py:790-817, holds).  The pipeline owns its device
 * workspace (two chunks of `chunk` structures, two streams); ps_host_inter_residue_geometry uploads, launches the
 * fused kernel and downloads chunk by chunk with copies and compute overlapped, and returns when every result
 * byte is in host memory.  xyz (B,L,A,3) f32, atom_mask (B,L,A) uint8 0/1, dist (B,L,L,A,A) f32, dist_mask
 * (B,L,L,A,A) uint8, omega/theta/phi (B,L,L) f32 — all HOST pointers, page-locked for full speed.
 */
int ps_host_pipeline_create(int chunk, int L, int A, void** pipeline);
int ps_host_pipeline_destroy(void* pipeline);
int ps_host_inter_residue_geometry(void* pipeline, const float* xyz, const uint8_t* atom_mask, int B,
                                   float* dist, uint8_t* dist_mask, float* omega, float* theta, float* phi);
int64_t ps_host_pipeline_launches(void* pipeline);  /* fused-kernel launches issued so far */

/*
 * K5 — one forward-diffusion step.  Replaces StructureBatch.diffuse_xyz
 * (protstruc/protstruc.py:864-878):
 *   out = fl( fl(sqrt(1-beta_b) * x) + fl(z * sqrt(beta_b)) )     (no FMA contraction)
 * noise != NULL: z is read from `noise` (bit-matches the reference given the same z).
 * noise == NULL: z ~ N(0,1) from Philox4x32-10 keyed by `seed`; with g = e + elem_offset the GLOBAL
 *   index of local element e, the element takes output lane (g & 3) of the counter (g >> 2, step), so
 *   the stream does not depend on the launch shape or on how the batch is sharded across GPUs
 *   (elem_offset = first global element of this shard; ANY value, not only multiples of 4).
 * x, out: `B * per_b` f32 elements (per_b = L*A*3); beta (B,) f32.  out may alias x.
 */
int ps_diffuse(const float* x, const float* beta, const float* noise,
               uint64_t seed, uint64_t step, uint64_t elem_offset,
               float* out, int B, int64_t per_b, void* stream);

/*
 * K5m — T consecutive diffusion steps fused in registers (one read, one write of x):
 *   for t in [0,T): x = fl(fl(sqrt(1-betas[t,b]) * x) + fl(z_t * sqrt(betas[t,b])))
 * z_t is the Philox stream of ps_diffuse with step = step0 + t, so the result is bit-identical
 * to T calls of ps_diffuse(noise = NULL).  betas (T,B) f32.
 */
int ps_diffuse_steps(const float* x, const float* betas, int T,
                     uint64_t seed, uint64_t step0, uint64_t elem_offset,
                     float* out, int B, int64_t per_b, void* stream);

/*
 * K5t — the same T steps with EVERY intermediate state kept: trajectory (T, B, per_b) f32, slice t = the state after
 * step t (what the reference's tutorial loop collects, docs/tutorials/diffusing_xyz_coordinates.ipynb).  T launches of
 * the one-step kernel issued back to back by the library (no per-step host round trip); bit-identical to T calls of
 * ps_diffuse.
 */
int ps_diffuse_trajectory(const float* x, const float* betas, int T,
                          uint64_t seed, uint64_t step0, uint64_t elem_offset,
                          float* trajectory, int B, int64_t per_b, void* stream);

/* Fills `out` with the N(0,1) stream ps_diffuse would use (for distribution tests). */
int ps_philox_normal(float* out, int64_t n, uint64_t seed, uint64_t step,
                     uint64_t elem_offset, void* stream);

/*
 * Free-function geometry on flat point lists, replacing protstruc/geometry.py:
 *   ps_geom_angle     — geometry.angle      (:39-71)   a,b,c  (n,3) → out (n)
 *   ps_geom_dihedral  — geometry.dihedral   (:74-124)  a,b,c,d (n,3) → out (n)
 *   ps_geom_gram_schmidt — geometry.gram_schmidt (:413-439) a,b,c (n,3) → out (n,3,3)
 * to_degree != 0 converts like torch.rad2deg / np.degrees (multiply by 180/pi in fp32).
 */
/* geometry.dot / norm / unit (protstruc/geometry.py:24-36) on n rows of D components: out (n), (n), (n, D). */
int ps_geom_dot(const float* x, const float* y, int64_t n, int D, float* out, void* stream);
int ps_geom_norm(const float* x, int64_t n, int D, float* out, void* stream);
int ps_geom_unit(const float* x, int64_t n, int D, float* out, void* stream);
int ps_geom_angle(const float* a, const float* b, const float* c, int64_t n,
                  int to_degree, float* out, void* stream);
int ps_geom_dihedral(const float* a, const float* b, const float* c, const float* d,
                     int64_t n, int to_degree, float* out, void* stream);
int ps_geom_gram_schmidt(const float* a, const float* b, const float* c, int64_t n,
                         float* out, void* stream);

/*
 * Tuning hook for K1 (benchmarks / profiling only; not part of the drop-in surface).
 * variant bit-field: bits 0-1 sqrt mode (0 = sqrt.approx.ftz.f32 [default], 1 = sqrt.approx.f32,
 * 2 = sqrt.rn.f32); bits 4-7 tile buffers per CTA override (0 = default); bit 8 = keep the staged kernel out
 * (any-A tile kernel); bit 9 = use the non-default number of warps per tile; bit 10 = diagnostic "stores only"
 * (no arithmetic, output content undefined: measures the memory-system ceiling of the kernel's write pattern);
 * bit 11 = force the lockstep schedule, bit 13 = force the cell schedule (default: chosen by length), bit 14 = lockstep
 * even with up to 35 % idle tile buffers; bit 12 = with bit 8: row kernel only (the fallback for atom counts whose tile does
 * not fit in shared memory); bits 16-23 = any-A tile kernel: pairs per tile in units of its alignment quantum
 * (0 = choose); bit 24 / 25 = any-A tile kernel: 128 / 256 threads per CTA; bit 15 = A = 15: the column-strip kernel of
 * round 1 instead of the linear-sweep kernel (comparison hook; the environment variable PROTSTRUC_B200_K1 = strip | sweep
 * does the same for a whole process, bit 27 = sweep regardless of it); bits 28-30 = linear-sweep kernel: the issuing lane
 * sleeps n x 100 ns after handing a tile to the TMA engine (pacing probe); bit 26 = linear-sweep kernel without its per-kind
 * pacing defaults (non-ftz square root for distances + byte mask, 400 ns for distances + fp32 mask).
 * bit 19 = fused call on the 5- / 10-atom layouts: keep ONE fused launch (default from 32 k pairs: distance tiles, then
 * the exact-sequence angle kernel — the fused tile kernel is angle-bound there), bit 20 = split any staged atom count
 * and size the same way (comparison hooks; identical bits either way).
 * L2 eviction policy of the bulk tile stores (0 = the launcher's default: evict_first for distances + byte mask on the
 * linear-sweep kernel, for the 10- / 14-atom strip kernels and for the any-A tile kernel with the
 * byte mask; none elsewhere): linear-sweep kernel bits 16-18 (1 = evict_first, 2 = evict_last, 3 / 4 = evict_first for
 * the distance / the mask tile only, 5 / 6 = streaming stores for the fused angle planes without / with evict_first
 * tiles, 7 = none); strip kernels bits 16-17 (1, 2, 3 = none); any-A tile kernel bits 28-29 (1, 2, 3 = none).  The
 * environment variable PROTSTRUC_B200_L2HINT (sweep-kernel coding) sets it for a whole process.  Results are
 * bit-identical whatever the policy.
 */
int ps_pair_dist_mask_ex(const float* xyz, const void* atom_mask, int mask_dtype,
                         float* dist, void* dist_mask,
                         int B, int L, int A, int variant, void* stream);
/*
 * What the most recent K1 launch of the calling host thread chose (tests assert on it that a shape took the path its
 * parity claim is about).  out[0..n) of: path (0 staged tile kernel, 1 any-A tile kernel, 2 row kernel), lock-step
 * schedule (0/1), CTAs, tile buffers of the grid, tile buffers taking part, strip stride in tiles, pairs per tile,
 * kernel launches of the call, linear-sweep kernel (1) or column-strip kernel (0).
 */
int ps_pair_dist_last_plan(int64_t* out, int n);
/* Diagnostic store ceiling: plain 128-bit stores of a non-uniform pattern over n floats (n % 4 == 0). */
int ps_debug_fill_pattern(float* out, int64_t n, int blocks_per_sm, void* stream);
/* Same hook for the fused kernel (ps_inter_residue_geometry with a variant bit-field). */
int ps_inter_residue_geometry_ex(const float* xyz, const void* atom_mask, int mask_dtype,
                                 float* dist, void* dist_mask,
                                 float* omega, float* theta, float* phi,
                                 int B, int L, int A, int variant, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PROTSTRUC_B200_H_ */
