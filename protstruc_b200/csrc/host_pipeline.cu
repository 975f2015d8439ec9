// Host-buffer entry to the full pairwise feature set: the call a user of the reference makes
// (StructureBatch.inter_residue_geometry on host tensors, protstruc/protstruc.py:790-817), as a native
// streaming pipeline.  Host arrays in, host arrays out; the library owns the device workspace here
// (created once per pipeline, never inside a kernel launch path).
//
// Structures are pushed through the GPU in chunks on two CUDA streams with double-buffered device
// workspaces, so the host->device copy, the fused kernel and the device->host copies of consecutive chunks
// overlap.  The device->host copy (298 MB per 512-residue structure) is what bounds this path: with pinned
// host buffers it runs at PCIe speed (~55 GB/s on the B200 boxes), the kernel itself is ~100x faster.

#include <new>

#include "common.cuh"

namespace ps {

int pair_dist_mask_impl(const float* xyz, const void* atom_mask, int mask_dtype, float* dist, void* dist_mask,
                        float* omega, float* theta, float* phi, int B, int L, int A, int variant,
                        cudaStream_t stream);

struct HostPipeline {
    static constexpr int kSlots = 2;
    int chunk = 0, L = 0, A = 0, device = 0;
    cudaStream_t streams[kSlots] = {nullptr, nullptr};
    float* xyz[kSlots] = {nullptr, nullptr};
    uint8_t* mask[kSlots] = {nullptr, nullptr};
    float* dist[kSlots] = {nullptr, nullptr};
    uint8_t* dist_mask[kSlots] = {nullptr, nullptr};
    float* angles[kSlots] = {nullptr, nullptr};  // omega | theta | phi
    long long launches = 0;
};

namespace {

void release(HostPipeline* p) {
    if (!p) return;
    int caller_device = 0;
    const bool switched = cudaGetDevice(&caller_device) == cudaSuccess && caller_device != p->device &&
                          cudaSetDevice(p->device) == cudaSuccess;
    for (int s = 0; s < HostPipeline::kSlots; ++s) {
        if (p->streams[s]) cudaStreamDestroy(p->streams[s]);
        cudaFree(p->xyz[s]);
        cudaFree(p->mask[s]);
        cudaFree(p->dist[s]);
        cudaFree(p->dist_mask[s]);
        cudaFree(p->angles[s]);
    }
    if (switched) cudaSetDevice(caller_device);
    delete p;
}

}  // namespace

int host_pipeline_create_impl(int chunk, int L, int A, HostPipeline** out) {
    PS_REQUIRE(out != nullptr, PS_ERR_NULL_POINTER, "host_pipeline_create: out is NULL");
    PS_REQUIRE(chunk > 0 && L > 0 && A >= 5, PS_ERR_BAD_SHAPE, "host_pipeline_create: chunk=%d L=%d A=%d", chunk, L, A);
    HostPipeline* p = new (std::nothrow) HostPipeline();
    PS_REQUIRE(p != nullptr, PS_ERR_CUDA, "host_pipeline_create: out of host memory");
    p->chunk = chunk;
    p->L = L;
    p->A = A;
    cudaError_t err = cudaGetDevice(&p->device);
    const size_t pairs = static_cast<size_t>(chunk) * L * L;
    const size_t elems = pairs * A * A;
    for (int s = 0; s < HostPipeline::kSlots && err == cudaSuccess; ++s) {
        err = cudaStreamCreateWithFlags(&p->streams[s], cudaStreamNonBlocking);
        if (err == cudaSuccess) err = cudaMalloc(&p->xyz[s], static_cast<size_t>(chunk) * L * A * 3 * sizeof(float));
        if (err == cudaSuccess) err = cudaMalloc(&p->mask[s], static_cast<size_t>(chunk) * L * A);
        if (err == cudaSuccess) err = cudaMalloc(&p->dist[s], elems * sizeof(float));
        if (err == cudaSuccess) err = cudaMalloc(&p->dist_mask[s], elems);
        if (err == cudaSuccess) err = cudaMalloc(&p->angles[s], 3 * pairs * sizeof(float));
    }
    if (err != cudaSuccess) {
        release(p);
        return cuda_fail(err, "host_pipeline_create");
    }
    *out = p;
    return PS_OK;
}

int host_pipeline_destroy_impl(HostPipeline* p) {
    release(p);
    return PS_OK;
}

int host_pipeline_run_impl(HostPipeline* p, const float* xyz_host, const uint8_t* mask_host, int B,
                           float* dist_host, uint8_t* dist_mask_host, float* omega_host, float* theta_host,
                           float* phi_host) {
    PS_REQUIRE(p != nullptr, PS_ERR_NULL_POINTER, "host_pipeline_run: pipeline is NULL");
    PS_REQUIRE(B > 0, PS_ERR_BAD_SHAPE, "host_pipeline_run: B=%d", B);
    PS_REQUIRE(xyz_host && mask_host && dist_host && dist_mask_host && omega_host && theta_host && phi_host,
               PS_ERR_NULL_POINTER, "host_pipeline_run: NULL host buffer");
    // The workspaces and streams belong to the device the pipeline was created on: run there whatever the caller's
    // current device is, and restore it on every way out.  On an error the copies already queued still target the
    // caller's host buffers, so both streams are drained before the error is returned.
    int caller_device = 0;
    cudaError_t dev_err = cudaGetDevice(&caller_device);
    if (dev_err == cudaSuccess && caller_device != p->device) dev_err = cudaSetDevice(p->device);
    if (dev_err != cudaSuccess) return cuda_fail(dev_err, "host_pipeline_run: select device");
    auto leave = [&](int rc) {
        if (rc != PS_OK)
            for (int s = 0; s < HostPipeline::kSlots; ++s) cudaStreamSynchronize(p->streams[s]);
        if (caller_device != p->device) cudaSetDevice(caller_device);
        return rc;
    };
    const int L = p->L, A = p->A;
    const size_t res_floats = static_cast<size_t>(L) * A * 3, res_mask = static_cast<size_t>(L) * A;
    const size_t pairs_per = static_cast<size_t>(L) * L, elems_per = pairs_per * A * A;
    int k = 0;
    for (int start = 0; start < B; start += p->chunk, ++k) {
        const int n = (B - start < p->chunk) ? B - start : p->chunk;
        const int s = k % HostPipeline::kSlots;
        cudaStream_t st = p->streams[s];
        cudaError_t err = cudaMemcpyAsync(p->xyz[s], xyz_host + start * res_floats, n * res_floats * sizeof(float),
                                          cudaMemcpyHostToDevice, st);
        if (err == cudaSuccess)
            err = cudaMemcpyAsync(p->mask[s], mask_host + start * res_mask, n * res_mask, cudaMemcpyHostToDevice, st);
        if (err != cudaSuccess) return leave(cuda_fail(err, "host_pipeline_run: host->device copy"));
        float* om = p->angles[s];
        float* th = om + static_cast<size_t>(p->chunk) * pairs_per;
        float* ph = th + static_cast<size_t>(p->chunk) * pairs_per;
        const int rc = pair_dist_mask_impl(p->xyz[s], p->mask[s], PS_MASK_BOOL, p->dist[s], p->dist_mask[s], om, th,
                                           ph, n, L, A, 0, st);
        if (rc != PS_OK) return leave(rc);
        ++p->launches;
        err = cudaMemcpyAsync(dist_host + start * elems_per, p->dist[s], n * elems_per * sizeof(float),
                              cudaMemcpyDeviceToHost, st);
        if (err == cudaSuccess)
            err = cudaMemcpyAsync(dist_mask_host + start * elems_per, p->dist_mask[s], n * elems_per,
                                  cudaMemcpyDeviceToHost, st);
        if (err == cudaSuccess)
            err = cudaMemcpyAsync(omega_host + start * pairs_per, om, n * pairs_per * sizeof(float),
                                  cudaMemcpyDeviceToHost, st);
        if (err == cudaSuccess)
            err = cudaMemcpyAsync(theta_host + start * pairs_per, th, n * pairs_per * sizeof(float),
                                  cudaMemcpyDeviceToHost, st);
        if (err == cudaSuccess)
            err = cudaMemcpyAsync(phi_host + start * pairs_per, ph, n * pairs_per * sizeof(float),
                                  cudaMemcpyDeviceToHost, st);
        if (err != cudaSuccess) return leave(cuda_fail(err, "host_pipeline_run: device->host copy"));
    }
    for (int s = 0; s < HostPipeline::kSlots; ++s) {
        const cudaError_t err = cudaStreamSynchronize(p->streams[s]);
        if (err != cudaSuccess) return leave(cuda_fail(err, "host_pipeline_run: synchronize"));
    }
    return leave(PS_OK);
}

long long host_pipeline_launches_impl(const HostPipeline* p) { return p ? p->launches : 0; }

}  // namespace ps
