// Minimal cluster probe for compute-sanitizer: every CTA of a cluster of 2 sends 1440 bytes of its shared memory to every CTA
// of the cluster (itself included) with cp.async.bulk.shared::cluster.shared::cta + a transaction mbarrier -- the exchange of
// udt_steps.cu reduced to its skeleton.  Prints the received data; under `compute-sanitizer --tool memcheck` it shows whether
// the tool accepts the pattern at all.
// nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/bulkprobe profiles/probes/bulkcopy_cluster_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned a, unsigned r) { unsigned o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__global__ void __cluster_dims__(2, 1, 1) probe(double* out)
{
    extern __shared__ __align__(16) double sm[];
    double* recv = sm;            // [2][180]
    double* send = sm + 360;      // [180]
    unsigned long long* bar = (unsigned long long*)(sm + 540);
    unsigned rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 180; i += blockDim.x) send[i] = 1000.0 * rank + i;
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(bar)), "r"(2 * 1440) : "memory");
    __syncthreads();
    if (tid < 32) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (tid < 2)
            asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(mapa(s32(recv + 180 * rank), tid)), "r"(s32(send)), "r"(1440), "r"(mapa(s32(bar), tid)) : "memory");
    }
    if (tid == 0) {
        asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" :: "r"(s32(bar)) : "memory");
    }
    __syncthreads();
    for (int i = tid; i < 360; i += blockDim.x) out[blockIdx.x * 360 + i] = recv[i];
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
int main()
{
    double* d; cudaMalloc(&d, 720 * 8);
    probe<<<2, 64, 541 * 8 + 8>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    double h[720]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int b = 0; b < 2; ++b) for (int r = 0; r < 2; ++r) for (int i = 0; i < 180; ++i) bad += h[b * 360 + r * 180 + i] != 1000.0 * r + i;
    printf("sync: %s, mismatches: %d\n", cudaGetErrorString(e), bad);
    return bad != 0;
}
