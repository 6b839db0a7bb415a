// ubench4 — can the SPG bucket hand-off live in L2?  (round-2 design measurements, scratch)
//   A  L2-resident streaming read / write / copy bandwidth vs footprint
//   B  HBM stream (evict-first) + ring write + ring read at the same time, ring footprint swept (resident vs not)
//   C  same with the ring touched in short unaligned runs (what a per-tile counting sort produces)
//   D  small TMA bulk copies shared->global and global->shared (issue rate per SM)
//   E  cross-SM flag round trip through L2 (st.release / ld.acquire)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__host__ __device__ inline uint64_t mix64(uint64_t x) { x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull; return x ^ (x >> 31); }

template <typename F> float timeit(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

// ---- A: MODE 0 read, 1 write, 2 copy (read first half, write second half) over `bytes`, `reps` passes
template <int MODE>
__global__ void __launch_bounds__(512) l2_stream(uint4* buf, size_t n16, int reps, unsigned long long* out) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
    const size_t half = n16 / 2;
    for (int r = 0; r < reps; r++) {
        if (MODE == 0) {
            for (size_t i = tid; i < n16; i += nt * 4) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) { size_t j = i + u * nt; v[u] = j < n16 ? __ldcg(buf + j) : make_uint4(0, 0, 0, 0); }
#pragma unroll
                for (int u = 0; u < 4; u++) { acc.x ^= v[u].x; acc.y += v[u].y; acc.z ^= v[u].z; acc.w += v[u].w; }
            }
        } else if (MODE == 1) {
            for (size_t i = tid; i < n16; i += nt) buf[i] = make_uint4((unsigned)i, r, 0, 1);
        } else {
            for (size_t i = tid; i < half; i += nt * 4) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) { size_t j = i + u * nt; v[u] = j < half ? __ldcg(buf + j) : make_uint4(0, 0, 0, 0); }
#pragma unroll
                for (int u = 0; u < 4; u++) { size_t j = i + u * nt; if (j < half) buf[half + j] = v[u]; }
            }
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 0x12345) atomicAdd(out, 1ull);
}

// ---- B / C: per thread-iteration: U 16-byte stream loads (evict-first), U ring stores, U ring loads.
// RUN = 0: ring touched in full coalesced lines; RUN > 0: each warp touches two runs of RUN rows (16 B each) at
// pseudo-random 16-byte aligned offsets (lanes beyond 2*RUN idle for the ring ops).
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 v;
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}
template <int RUN, bool DO_STREAM, bool DO_WR, bool DO_RD>
__global__ void __launch_bounds__(512) ring_mix(const uint4* __restrict__ stream, size_t n16, uint4* ring, size_t ring16, unsigned long long* out) {
    constexpr int U = 4;
    uint4 acc = make_uint4(0, 0, 0, 0);
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
    const unsigned lane = threadIdx.x & 31;
    const size_t warp = tid >> 5;
    size_t it = 0;
    for (size_t i = tid; i + (U - 1) * nt < n16; i += nt * U, it++) {
        uint4 v[U], w[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = DO_STREAM ? ld_stream(stream + i + u * nt) : make_uint4((unsigned)i, u, 2, 3);
        size_t wo[U], ro[U];
        bool act = true;
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (RUN == 0) {
                wo[u] = (i + u * nt) % ring16;
                ro[u] = (i + u * nt + ring16 / 2 + 12345 * 8) % ring16;
            } else {
                const unsigned half = lane / RUN;  // 0 or 1 (or more = idle)
                act = half < 2;
                const uint64_t h = mix64((warp * 1315423911ull + it) * 8 + u * 2 + (half & 1));
                wo[u] = (h % (ring16 - RUN)) + lane % RUN;
                ro[u] = ((h >> 20) % (ring16 - RUN)) + lane % RUN;
            }
        }
        if (DO_RD) {
#pragma unroll
            for (int u = 0; u < U; u++) w[u] = act ? __ldcg(ring + ro[u]) : make_uint4(0, 0, 0, 0);
        }
        if (DO_WR) {
#pragma unroll
            for (int u = 0; u < U; u++) if (act) ring[wo[u]] = v[u];
        } else {
#pragma unroll
            for (int u = 0; u < U; u++) { acc.x ^= v[u].x; acc.y += v[u].y; }
        }
        if (DO_RD) {
#pragma unroll
            for (int u = 0; u < U; u++) { acc.z ^= w[u].x; acc.w += w[u].y; }
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 0x12345) atomicAdd(out, 1ull);
}

// ---- D: small bulk copies.  DIR 0: shared -> global (bulk_group), DIR 1: global -> shared (mbarrier)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int DIR>
__global__ void __launch_bounds__(128) bulk_small(uint4* region, size_t region16, int bytes, int ops_per_thread, int issuers, unsigned long long* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t mbar[4];
    const int tid = threadIdx.x;
    if (tid < 4) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar[tid])), "r"(1) : "memory");
    for (int i = tid; i < 8192; i += 128) ((uint32_t*)sm)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int n16 = bytes / 16;
    if ((tid & 31) == 0 && (tid >> 5) < issuers) {
        const int w = tid >> 5;
        uint32_t phase = 0;
        for (int k = 0; k < ops_per_thread; k += 8) {
            if (DIR == 1) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar[w])), "r"(8 * bytes) : "memory");
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint64_t h = mix64(((uint64_t)blockIdx.x * 4 + w) * 1000003ull + k + j);
                uint4* g = region + (h % (region16 - n16));
                unsigned char* s = sm + w * 8192 + j * 1024;
                if (DIR == 0) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smem_u32(s)), "r"(bytes) : "memory");
                else asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(s)), "l"(g), "r"(bytes), "r"(smem_u32(&mbar[w])) : "memory");
            }
            if (DIR == 0) {
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            } else {
                uint32_t ok = 0;
                while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&mbar[w])), "r"(phase) : "memory");
                phase ^= 1;
            }
        }
        if (DIR == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (((uint32_t*)sm)[tid] == 0xdeadbeef) atomicAdd(out, 1ull);
}

// ---- E: flag ping-pong between CTA 0 and CTA `peer`
__global__ void flag_pingpong(volatile unsigned int* flags, int peer, int rounds, long long* cycles) {
    if (threadIdx.x != 0) return;
    unsigned int* f0 = (unsigned int*)flags, *f1 = (unsigned int*)flags + 64;
    if (blockIdx.x == 0) {
        long long t0 = clock64();
        for (int r = 1; r <= rounds; r++) {
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(f0), "r"(r) : "memory");
            unsigned int v = 0;
            while (v != (unsigned)r) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f1) : "memory");
        }
        *cycles = clock64() - t0;
    } else if ((int)blockIdx.x == peer) {
        for (int r = 1; r <= rounds; r++) {
            unsigned int v = 0;
            while (v != (unsigned)r) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f0) : "memory");
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(f1), "r"(r) : "memory");
        }
    }
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, L2 %d MB, persisting max %d MB, clock %d MHz\n", prop.name, sms, prop.l2CacheSize >> 20, prop.persistingL2CacheMaxSize >> 20, prop.clockRate / 1000);
    unsigned long long* d_out; CK(cudaMalloc(&d_out, 64)); CK(cudaMemset(d_out, 0, 64));
    const size_t stream_bytes = 4ull << 30;
    uint4* d_stream; CK(cudaMalloc(&d_stream, stream_bytes)); CK(cudaMemset(d_stream, 1, stream_bytes));
    const size_t ring_max = 2ull << 30;
    uint4* d_ring; CK(cudaMalloc(&d_ring, ring_max)); CK(cudaMemset(d_ring, 2, ring_max));
    const int grid = sms * 4;

    printf("\n== A: streaming over a footprint, GB/s (read / write / copy r+w) ==\n");
    for (size_t mb : {4, 8, 16, 24, 32, 48, 64, 96, 128, 256, 1024}) {
        size_t bytes = mb << 20, n16 = bytes / 16;
        int reps = (int)((8ull << 30) / bytes); if (reps < 2) reps = 2;
        float tr = timeit([&] { l2_stream<0><<<grid, 512>>>(d_ring, n16, reps, d_out); });
        float tw = timeit([&] { l2_stream<1><<<grid, 512>>>(d_ring, n16, reps, d_out); });
        float tc = timeit([&] { l2_stream<2><<<grid, 512>>>(d_ring, n16, reps, d_out); });
        printf("footprint %5zu MB: read %7.0f  write %7.0f  copy(r+w bytes) %7.0f GB/s\n", mb, bytes * (double)reps / tr / 1e6, bytes * (double)reps / tw / 1e6, bytes * (double)reps / tc / 1e6);
    }

    printf("\n== B: stream 4 GB (16 B units) + ring write + ring read per unit; Gunits/s (1 unit = one 16-byte row) ==\n");
    const size_t n16 = stream_bytes / 16;
    {
        float t = timeit([&] { ring_mix<0, true, false, false><<<grid, 512>>>(d_stream, n16, d_ring, 1 << 20, d_out); });
        printf("stream only                 : %6.1f Gunits/s (%5.0f GB/s)\n", n16 / t / 1e6, stream_bytes / t / 1e6);
    }
    for (size_t mb : {8, 16, 24, 32, 48, 64, 2048}) {
        size_t r16 = (mb << 20) / 16;
        float t1 = timeit([&] { ring_mix<0, true, true, true><<<grid, 512>>>(d_stream, n16, d_ring, r16, d_out); });
        float t2 = timeit([&] { ring_mix<0, true, true, false><<<grid, 512>>>(d_stream, n16, d_ring, r16, d_out); });
        float t3 = timeit([&] { ring_mix<0, false, true, true><<<grid, 512>>>(d_stream, n16, d_ring, r16, d_out); });
        printf("ring %5zu MB lines         : stream+wr+rd %6.1f   stream+wr %6.1f   wr+rd (no stream) %6.1f Gunits/s\n", mb, n16 / t1 / 1e6, n16 / t2 / 1e6, n16 / t3 / 1e6);
    }
    printf("\n== C: ring touched in runs (two runs of RUN rows per warp op); Gunits/s of ring rows ==\n");
    for (size_t mb : {16, 32, 48, 2048}) {
        size_t r16 = (mb << 20) / 16;
        // per iteration a warp moves 32 stream units but only 2*RUN ring rows; report ring rows/s and stream units/s
        float t14 = timeit([&] { ring_mix<14, true, true, true><<<grid, 512>>>(d_stream, n16, d_ring, r16, d_out); });
        float t7 = timeit([&] { ring_mix<7, true, true, true><<<grid, 512>>>(d_stream, n16, d_ring, r16, d_out); });
        float t14n = timeit([&] { ring_mix<14, false, true, true><<<grid, 512>>>(d_stream, n16, d_ring, r16, d_out); });
        float t16n = timeit([&] { ring_mix<16, false, true, true><<<grid, 512>>>(d_stream, n16, d_ring, r16, d_out); });
        printf("ring %5zu MB: RUN14 stream %6.1f ring %6.1f | RUN7 stream %6.1f ring %6.1f | no stream: RUN14 ring %6.1f  RUN16 ring %6.1f Gunits/s\n", mb,
               n16 / t14 / 1e6, n16 * (28.0 / 32) / t14 / 1e6, n16 / t7 / 1e6, n16 * (14.0 / 32) / t7 / 1e6, n16 * (28.0 / 32) / t14n / 1e6, n16 / t16n / 1e6);
    }

    printf("\n== D: small TMA bulk copies, 1 CTA/SM x issuers warps; Mops/s per SM, GB/s total ==\n");
    CK(cudaFuncSetAttribute(bulk_small<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(bulk_small<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int bytes : {128, 224, 448, 896, 1024}) {
        for (int issuers : {1, 4}) {
            const int ops = 4096;
            size_t r16 = (32ull << 20) / 16;
            float t0 = timeit([&] { bulk_small<0><<<sms, 128, 32768>>>(d_ring, r16, bytes, ops, issuers, d_out); });
            float t1 = timeit([&] { bulk_small<1><<<sms, 128, 32768>>>(d_ring, r16, bytes, ops, issuers, d_out); });
            double tot = (double)sms * issuers * ops;
            printf("bytes %4d issuers %d: S2G %6.2f Mops/s/SM %7.0f GB/s | G2S %6.2f Mops/s/SM %7.0f GB/s\n", bytes, issuers,
                   ops * issuers / t0 / 1e3, tot * bytes / t0 / 1e6, ops * issuers / t1 / 1e3, tot * bytes / t1 / 1e6);
        }
    }

    printf("\n== E: flag ping-pong through L2 (release/acquire, gpu scope), cycles per round trip ==\n");
    unsigned int* d_flags; CK(cudaMalloc(&d_flags, 1024)); long long* d_cyc; CK(cudaMalloc(&d_cyc, 8));
    for (int peer : {1, 2, 37, 74, 100, 147}) {
        CK(cudaMemset(d_flags, 0, 1024));
        flag_pingpong<<<sms, 32>>>(d_flags, peer, 1000, d_cyc);
        CK(cudaDeviceSynchronize());
        long long c; CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
        printf("CTA 0 <-> CTA %3d: %lld cycles per round trip (two one-way hand-offs)\n", peer, c / 1000);
    }
    return 0;
}
