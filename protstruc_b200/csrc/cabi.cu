// extern "C" boundary of libprotstruc_b200.so — declarations in include/protstruc_b200.h.
// Thin argument adapters over the *_impl launchers; no device work happens here.

#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace ps {

// launchers defined in the kernel translation units
int pair_dist_mask_impl(const float*, const void*, int, float*, void*, float*, float*, float*, int,
                        int, int, int, cudaStream_t);
int pair_dist_mask_compact_impl(const float*, const void*, int, float*, void*, float*, float*, float*, float*,
                                float*, float*, int, int, int, int, cudaStream_t);
int pair_angles_impl(const float*, int, int, int, const int*, int, const int*, int, int, float*,
                     cudaStream_t);
int pair_angles_variant_impl(const float*, int, int, int, const int*, int, const int*, int, int, float*, int,
                             cudaStream_t);
int trrosetta_angles_impl(const float*, int, int, int, int, float*, float*, float*, cudaStream_t);
int trrosetta_angles_variant_impl(const float*, int, int, int, int, float*, float*, float*, int, cudaStream_t);
int backbone_impl(const float*, const uint8_t*, const float*, int, int, int, int, int, int, float*,
                  uint8_t*, float*, cudaStream_t);
int geom_angle_impl(const float*, const float*, const float*, long long, int, float*, cudaStream_t);
int geom_dihedral_impl(const float*, const float*, const float*, const float*, long long, int,
                       float*, cudaStream_t);
int geom_gram_schmidt_impl(const float*, const float*, const float*, long long, float*,
                           cudaStream_t);
int geom_rowwise_impl(const float*, const float*, long long, int, int, float*, cudaStream_t);
int masked_stats_impl(const float*, const void*, int, int, int, int, float*, float*, float*,
                      cudaStream_t);
int masked_stats_variant_impl(const float*, const void*, int, int, int, int, float*, float*, float*, int,
                              cudaStream_t);
int scale_shift_impl(const float*, const float*, const float*, int, int, int, float*, cudaStream_t);
int translate_impl(const float*, const float*, int, int, int, int, float*, cudaStream_t);
int center_of_mass_impl(const float*, int, int, int, int, float*, cudaStream_t);
int diffuse_impl(const float*, const float*, int, const float*, uint64_t, uint64_t, uint64_t, float*,
                 int, long long, cudaStream_t);
int diffuse_trajectory_impl(const float*, const float*, int, uint64_t, uint64_t, uint64_t, float*, int, long long,
                            cudaStream_t);
int philox_normal_impl(float*, long long, uint64_t, uint64_t, uint64_t, cudaStream_t);
int kabsch_impl(const float*, const float*, const uint8_t*, int, int, int, float*, float*, cudaStream_t);
int topk_nearest_impl(const float*, const uint8_t*, const float*, int, int, int, int, int, float*, uint8_t*,
                      cudaStream_t);
int debug_fill_pattern_impl(float*, long long, int, cudaStream_t);
int pair_dist_last_plan_impl(long long*, int);
void pair_sweep_set_push_target(void* const*, int, void*, long long, long long);
int pair_sweep_max_peers();
bool pair_sweep_supported(const float*, const void*, int, const float*, const void*, int, int);
struct HostPipeline;
int host_pipeline_create_impl(int, int, int, HostPipeline**);
int host_pipeline_destroy_impl(HostPipeline*);
int host_pipeline_run_impl(HostPipeline*, const float*, const uint8_t*, int, float*, uint8_t*, float*, float*, float*);
long long host_pipeline_launches_impl(const HostPipeline*);
int host_pdb_parse_impl(const char*, long long, int, float*, uint8_t*, int32_t*, char*, int32_t*, char*, char*, int*);
int local_xyz_impl(const float*, int, int, int, int, int, int, int, float*, cudaStream_t);
int rotate_impl(const float*, const float*, int, int, int, int, float*, cudaStream_t);
int frames_to_backbone_impl(const float*, const float*, const float*, int, int, int, int, float*, float*,
                            cudaStream_t);
int translate_bcast_impl(const float*, const float*, long long, long long, long long, int, int, int,
                         float*, cudaStream_t);

namespace {
thread_local char g_last_error[512] = "";
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t err, const char* what) {
    set_error("%s: %s (%s)", what, cudaGetErrorString(err), cudaGetErrorName(err));
    return PS_ERR_CUDA;
}

int check_launch(const char* kernel_name) {
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return cuda_fail(err, kernel_name);
    return PS_OK;
}

namespace {
int g_reserved_sms = 0;  // ps_reserve_sms: SMs the persistent kernels leave to concurrently running work
}

int sm_count_for_current_device() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_sms = 0;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return cuda_fail(err, "cudaGetDevice");
    if (dev != cached_dev) {
        int sms = 0;
        err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (err != cudaSuccess) return cuda_fail(err, "cudaDeviceGetAttribute(SM count)");
        cached_dev = dev;
        cached_sms = sms;
    }
    const int usable = cached_sms - g_reserved_sms;
    return usable > 0 ? usable : 1;
}

}  // namespace ps

#define PS_STREAM(s) (static_cast<cudaStream_t>(s))
#define PS_STRINGIFY_(x) #x
#define PS_STRINGIFY(x) PS_STRINGIFY_(x)

extern "C" {

int ps_abi_version(void) { return 1; }

const char* ps_build_info(void) {
    return "protstruc_b200 sm_100a; nvcc " PS_STRINGIFY(__CUDACC_VER_MAJOR__) "." PS_STRINGIFY(
        __CUDACC_VER_MINOR__) "." PS_STRINGIFY(__CUDACC_VER_BUILD__);
}

const char* ps_last_error_string(void) { return ps::g_last_error; }

int ps_reserve_sms(int n) {
    if (n < 0) {
        ps::set_error("ps_reserve_sms: n=%d must be >= 0", n);
        return PS_ERR_BAD_SHAPE;
    }
    const int before = ps::g_reserved_sms;
    ps::g_reserved_sms = n;
    return before;
}

int ps_device_sm_count(int device) {
    int sms = 0;
    const cudaError_t err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (err != cudaSuccess) return ps::cuda_fail(err, "cudaDeviceGetAttribute(SM count)");
    return sms;
}

int ps_pair_dist_mask(const float* xyz, const void* atom_mask, int mask_dtype, float* dist,
                      void* dist_mask, int B, int L, int A, void* stream) {
    return ps::pair_dist_mask_impl(xyz, atom_mask, mask_dtype, dist, dist_mask, nullptr, nullptr,
                                   nullptr, B, L, A, 0, PS_STREAM(stream));
}

int ps_pair_dist_mask_ex(const float* xyz, const void* atom_mask, int mask_dtype, float* dist,
                         void* dist_mask, int B, int L, int A, int variant, void* stream) {
    return ps::pair_dist_mask_impl(xyz, atom_mask, mask_dtype, dist, dist_mask, nullptr, nullptr,
                                   nullptr, B, L, A, variant, PS_STREAM(stream));
}

int ps_pair_angles(const float* xyz, int B, int L, int A, const int* slots_i, int n_i,
                   const int* slots_j, int n_j, int kind, float* out, void* stream) {
    return ps::pair_angles_impl(xyz, B, L, A, slots_i, n_i, slots_j, n_j, kind, out,
                                PS_STREAM(stream));
}

int ps_pair_angles_ex(const float* xyz, int B, int L, int A, const int* slots_i, int n_i,
                      const int* slots_j, int n_j, int kind, float* out, int variant, void* stream) {
    return ps::pair_angles_variant_impl(xyz, B, L, A, slots_i, n_i, slots_j, n_j, kind, out, variant,
                                        PS_STREAM(stream));
}

int ps_trrosetta_angles(const float* xyz, int B, int L, int A, int virtual_cb, float* omega,
                        float* theta, float* phi, void* stream) {
    return ps::trrosetta_angles_impl(xyz, B, L, A, virtual_cb, omega, theta, phi, PS_STREAM(stream));
}

int ps_trrosetta_angles_ex(const float* xyz, int B, int L, int A, int virtual_cb, float* omega,
                           float* theta, float* phi, int variant, void* stream) {
    return ps::trrosetta_angles_variant_impl(xyz, B, L, A, virtual_cb, omega, theta, phi, variant, PS_STREAM(stream));
}

int ps_inter_residue_geometry(const float* xyz, const void* atom_mask, int mask_dtype, float* dist,
                              void* dist_mask, float* omega, float* theta, float* phi, int B, int L,
                              int A, void* stream) {
    if (!(omega && theta && phi && dist && dist_mask && atom_mask)) {
        ps::set_error("inter_residue_geometry: all inputs and outputs are required");
        return PS_ERR_NULL_POINTER;
    }
    // staged atom counts (5, 10, 14, 15) with L >= pairs per tile: one fused launch; otherwise the any-A tile
    // kernel followed by the fused angle kernel (decided inside pair_dist_mask_impl)
    return ps::pair_dist_mask_impl(xyz, atom_mask, mask_dtype, dist, dist_mask, omega, theta, phi, B,
                                   L, A, 0, PS_STREAM(stream));
}

int ps_inter_residue_geometry_compact(const float* xyz, const void* atom_mask, int mask_dtype, float* dist,
                                      void* dist_mask, float* compact, int B, int L, int A, void* stream) {
    if (!(compact && dist && dist_mask && atom_mask)) {
        ps::set_error("inter_residue_geometry_compact: all inputs and outputs are required");
        return PS_ERR_NULL_POINTER;
    }
    if (B <= 0 || L <= 0) {
        ps::set_error("inter_residue_geometry_compact: B=%d L=%d must be > 0", B, L);
        return PS_ERR_BAD_SHAPE;
    }
    const long long plane = static_cast<long long>(B) * L * L;
    return ps::pair_dist_mask_compact_impl(xyz, atom_mask, mask_dtype, dist, dist_mask, compact, compact + plane,
                                           compact + 2 * plane, compact + 3 * plane, compact + 4 * plane,
                                           compact + 5 * plane, B, L, A, 0, PS_STREAM(stream));
}

int ps_inter_residue_geometry_push(const float* xyz, const void* atom_mask, int mask_dtype, float* dist,
                                   void* dist_mask, void* const* peer_buffers, int world, int rank,
                                   void* multicast_buffer, int shard, int B, int L, int A, void* stream) {
    if (!(dist && dist_mask && atom_mask && peer_buffers)) {
        ps::set_error("inter_residue_geometry_push: all inputs and outputs are required");
        return PS_ERR_NULL_POINTER;
    }
    if (world < 1 || world > ps::pair_sweep_max_peers() || rank < 0 || rank >= world || B < 1 || B > shard || L < 1) {
        ps::set_error("inter_residue_geometry_push: world=%d (1..%d) rank=%d B=%d shard=%d L=%d", world,
                      ps::pair_sweep_max_peers(), rank, B, shard, L);
        return PS_ERR_BAD_SHAPE;
    }
    for (int r = 0; r < world; ++r) {
        if (peer_buffers[r] == nullptr) {
            ps::set_error("inter_residue_geometry_push: peer buffer %d is NULL", r);
            return PS_ERR_NULL_POINTER;
        }
    }
    if (!ps::pair_sweep_supported(xyz, atom_mask, mask_dtype, dist, dist_mask, L, A)) {
        ps::set_error("inter_residue_geometry_push: needs the linear-sweep kernel (A = 15, L >= 32, 16-byte aligned "
                      "arrays); use ps_inter_residue_geometry_compact + an NCCL all-gather for other shapes");
        return PS_ERR_BAD_SHAPE;
    }
    const long long slab = static_cast<long long>(shard) * L * L;
    ps::pair_sweep_set_push_target(peer_buffers, world, multicast_buffer, slab * world, slab * rank);
    // variant bit 27: the linear-sweep kernel regardless of the PROTSTRUC_B200_K1 environment override; bit 26: no pacing
    // defaults (they are chosen for the UNFUSED kinds, this launch evaluates the angles)
    const int rc = ps::pair_dist_mask_impl(xyz, atom_mask, mask_dtype, dist, dist_mask, nullptr, nullptr, nullptr, B, L,
                                           A, (1 << 27) | (1 << 26), PS_STREAM(stream));
    // the sweep launcher consumes the target; if the call failed before reaching it, no later launch of this thread
    // may inherit the peer pointers
    ps::pair_sweep_set_push_target(nullptr, 0, nullptr, 0, 0);
    return rc;
}

int ps_inter_residue_geometry_ex(const float* xyz, const void* atom_mask, int mask_dtype, float* dist,
                                 void* dist_mask, float* omega, float* theta, float* phi, int B,
                                 int L, int A, int variant, void* stream) {
    if (!(omega && theta && phi && dist && dist_mask && atom_mask)) {
        ps::set_error("inter_residue_geometry_ex: all inputs and outputs are required");
        return PS_ERR_NULL_POINTER;
    }
    return ps::pair_dist_mask_impl(xyz, atom_mask, mask_dtype, dist, dist_mask, omega, theta, phi, B,
                                   L, A, variant, PS_STREAM(stream));
}

int ps_backbone(const float* xyz, const uint8_t* residue_mask, const float* chain_idx, int B, int L,
                int A, int a1, int a2, int a3, float* dihedrals, uint8_t* dihedral_mask,
                float* frames, void* stream) {
    return ps::backbone_impl(xyz, residue_mask, chain_idx, B, L, A, a1, a2, a3, dihedrals,
                             dihedral_mask, frames, PS_STREAM(stream));
}

int ps_masked_stats(const float* xyz, const void* atom_mask, int mask_dtype, int B, int L, int A,
                    float* mu, float* sd, float* xyz_out, void* stream) {
    return ps::masked_stats_impl(xyz, atom_mask, mask_dtype, B, L, A, mu, sd, xyz_out,
                                 PS_STREAM(stream));
}

int ps_masked_stats_ex(const float* xyz, const void* atom_mask, int mask_dtype, int B, int L, int A,
                       float* mu, float* sd, float* xyz_out, int variant, void* stream) {
    return ps::masked_stats_variant_impl(xyz, atom_mask, mask_dtype, B, L, A, mu, sd, xyz_out, variant,
                                         PS_STREAM(stream));
}

int ps_scale_shift(const float* xyz, const float* scale, const float* shift, int B, int L, int A,
                   float* xyz_out, void* stream) {
    return ps::scale_shift_impl(xyz, scale, shift, B, L, A, xyz_out, PS_STREAM(stream));
}

int ps_center_of_mass(const float* xyz, int B, int L, int A, int slot, float* out, void* stream) {
    return ps::center_of_mass_impl(xyz, B, L, A, slot, out, PS_STREAM(stream));
}

int ps_translate(const float* xyz, const float* t, int t_rows, int B, int L, int A, float* xyz_out,
                 void* stream) {
    return ps::translate_impl(xyz, t, t_rows, B, L, A, xyz_out, PS_STREAM(stream));
}

int ps_local_xyz(const float* xyz, int B, int L, int A, int a1, int a2, int a3, int ca_slot,
                 float* out, void* stream) {
    return ps::local_xyz_impl(xyz, B, L, A, a1, a2, a3, ca_slot, out, PS_STREAM(stream));
}

int ps_rotate(const float* xyz, const float* rotation, int rot_rows, int B, int L, int A,
              float* xyz_out, void* stream) {
    return ps::rotate_impl(xyz, rotation, rot_rows, B, L, A, xyz_out, PS_STREAM(stream));
}

int ps_frames_to_backbone(const float* orientations, const float* translations, const float* ideal,
                          int n_ideal, int B, int L, int A, float* xyz, float* atom_mask,
                          void* stream) {
    return ps::frames_to_backbone_impl(orientations, translations, ideal, n_ideal, B, L, A, xyz,
                                       atom_mask, PS_STREAM(stream));
}

int ps_translate_bcast(const float* xyz, const float* t, int64_t stride_b, int64_t stride_l,
                       int64_t stride_a, int B, int L, int A, float* xyz_out, void* stream) {
    return ps::translate_bcast_impl(xyz, t, stride_b, stride_l, stride_a, B, L, A, xyz_out,
                                    PS_STREAM(stream));
}

int ps_kabsch(const float* source, const float* target, const uint8_t* mask, int target_rows, int B,
              int n_atoms, float* rotation, float* translation, void* stream) {
    return ps::kabsch_impl(source, target, mask, target_rows, B, n_atoms, rotation, translation,
                           PS_STREAM(stream));
}

int ps_topk_nearest_residue_mask(const float* xyz, const uint8_t* valid, const float* query, int n_query,
                                 int L, int A, int ca_slot, int k, float* scratch, uint8_t* out,
                                 void* stream) {
    return ps::topk_nearest_impl(xyz, valid, query, n_query, L, A, ca_slot, k, scratch, out,
                                 PS_STREAM(stream));
}

int ps_host_pdb_parse(const char* text, int64_t len, int capacity, float* xyz, uint8_t* atom_mask,
                      int32_t* chain_idx, char* chain_id, int32_t* residue_number, char* insertion_code,
                      char* one_letter, int* n_residues) {
    return ps::host_pdb_parse_impl(text, len, capacity, xyz, atom_mask, chain_idx, chain_id, residue_number,
                                   insertion_code, one_letter, n_residues);
}

int ps_host_pipeline_create(int chunk, int L, int A, void** pipeline) {
    return ps::host_pipeline_create_impl(chunk, L, A, reinterpret_cast<ps::HostPipeline**>(pipeline));
}

int ps_host_pipeline_destroy(void* pipeline) {
    return ps::host_pipeline_destroy_impl(static_cast<ps::HostPipeline*>(pipeline));
}

int ps_host_inter_residue_geometry(void* pipeline, const float* xyz, const uint8_t* atom_mask, int B,
                                   float* dist, uint8_t* dist_mask, float* omega, float* theta, float* phi) {
    return ps::host_pipeline_run_impl(static_cast<ps::HostPipeline*>(pipeline), xyz, atom_mask, B, dist, dist_mask,
                                      omega, theta, phi);
}

int64_t ps_host_pipeline_launches(void* pipeline) {
    return ps::host_pipeline_launches_impl(static_cast<const ps::HostPipeline*>(pipeline));
}

int ps_pair_dist_last_plan(int64_t* out, int n) {
    return ps::pair_dist_last_plan_impl(reinterpret_cast<long long*>(out), n);
}

int ps_debug_fill_pattern(float* out, int64_t n, int blocks_per_sm, void* stream) {
    return ps::debug_fill_pattern_impl(out, n, blocks_per_sm, PS_STREAM(stream));
}

int ps_diffuse(const float* x, const float* beta, const float* noise, uint64_t seed, uint64_t step,
               uint64_t elem_offset, float* out, int B, int64_t per_b, void* stream) {
    return ps::diffuse_impl(x, beta, 1, noise, seed, step, elem_offset, out, B, per_b,
                            PS_STREAM(stream));
}

int ps_diffuse_steps(const float* x, const float* betas, int T, uint64_t seed, uint64_t step0,
                     uint64_t elem_offset, float* out, int B, int64_t per_b, void* stream) {
    return ps::diffuse_impl(x, betas, T, nullptr, seed, step0, elem_offset, out, B, per_b,
                            PS_STREAM(stream));
}

int ps_diffuse_trajectory(const float* x, const float* betas, int T, uint64_t seed, uint64_t step0,
                          uint64_t elem_offset, float* trajectory, int B, int64_t per_b, void* stream) {
    return ps::diffuse_trajectory_impl(x, betas, T, seed, step0, elem_offset, trajectory, B, per_b, PS_STREAM(stream));
}

int ps_philox_normal(float* out, int64_t n, uint64_t seed, uint64_t step, uint64_t elem_offset,
                     void* stream) {
    return ps::philox_normal_impl(out, n, seed, step, elem_offset, PS_STREAM(stream));
}

int ps_geom_angle(const float* a, const float* b, const float* c, int64_t n, int to_degree,
                  float* out, void* stream) {
    return ps::geom_angle_impl(a, b, c, n, to_degree, out, PS_STREAM(stream));
}

int ps_geom_dihedral(const float* a, const float* b, const float* c, const float* d, int64_t n,
                     int to_degree, float* out, void* stream) {
    return ps::geom_dihedral_impl(a, b, c, d, n, to_degree, out, PS_STREAM(stream));
}

int ps_geom_dot(const float* x, const float* y, int64_t n, int D, float* out, void* stream) {
    return ps::geom_rowwise_impl(x, y, n, D, 0, out, PS_STREAM(stream));
}

int ps_geom_norm(const float* x, int64_t n, int D, float* out, void* stream) {
    return ps::geom_rowwise_impl(x, nullptr, n, D, 1, out, PS_STREAM(stream));
}

int ps_geom_unit(const float* x, int64_t n, int D, float* out, void* stream) {
    return ps::geom_rowwise_impl(x, nullptr, n, D, 2, out, PS_STREAM(stream));
}

int ps_geom_gram_schmidt(const float* a, const float* b, const float* c, int64_t n, float* out,
                         void* stream) {
    return ps::geom_gram_schmidt_impl(a, b, c, n, out, PS_STREAM(stream));
}

}  // extern "C"
