// Developer microbenchmark (not part of libqsb): how fast can one B200 stream a state-sized vector through
// the shared-memory tile ring that hpsi_ws_kernel uses?  Build + run on the GPU box:
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/stream_bench tools/stream_bench.cu
//   /tmp/stream_bench
//
// Cases: (1) plain LDG.128 grid-stride read (HBM read peak reference), (2) read + write (copy),
// (3) TMA ring: persistent CTA per SM, one producer warp, S stages of T bytes, consumers touch the tile and
// release it; tile -> CTA assignment static (round robin) or by a global ticket counter.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../qse_b200/csrc/common.cuh"

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            printf("%s failed: %s\n", #x, cudaGetErrorString(e_));                 \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

__global__ void __launch_bounds__(256) ldg_read(const double2* __restrict__ x, u64 n, double* out) {
    double acc = 0.0;
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (u64)gridDim.x * blockDim.x) {
        const double2 v = ldg2_volatile(x + k);
        acc += v.x + v.y;
    }
    if (acc == 1.2345e300) out[0] = acc;
}
__global__ void __launch_bounds__(256) ldg_copy(const double2* __restrict__ x, double2* __restrict__ y, u64 n) {
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (u64)gridDim.x * blockDim.x) y[k] = ldg2_volatile(x + k);
}

struct RingArgs {
    const double2* x;
    double2* y;     // write back the tile (consumer stores) when non-null
    u64 n_tiles;
    int tile_amps;  // amplitudes per tile
    int stages;
    int copies;     // bulk copies per tile
    int use_ticket;
    unsigned int* ticket;
    int lds_per_thread;  // LDS.128 each consumer thread performs per 8 amplitudes it owns (>= 1)
    double* out;
};

constexpr int kConsumers = 16;
constexpr int kThreadsRing = (kConsumers + 1) * 32;

__global__ void __launch_bounds__(kThreadsRing, 1) ring_kernel(const RingArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* tiles = reinterpret_cast<double2*>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[16], empty_bar[16];
    __shared__ long long item_tile[16];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (t == 0) {
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumers);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const u64 tile_bytes = (u64)a.tile_amps * sizeof(double2);
    if (warp == kConsumers) {
        unsigned int pf = 0;
        if (a.use_ticket && lane == 0) pf = atomicAdd(a.ticket, 1u);
        u64 next_static = blockIdx.x;
        for (u64 it = 0;; ++it) {
            const int s = (int)(it % a.stages);
            const u64 use = it / a.stages;
            long long tile;
            if (a.use_ticket) {
                const u64 b = (u64)__shfl_sync(0xffffffffu, pf, 0);
                tile = b < a.n_tiles ? (long long)b : -1;
                if (tile >= 0 && lane == 0) pf = atomicAdd(a.ticket, 1u);
            } else {
                tile = next_static < a.n_tiles ? (long long)next_static : -1;
                next_static += gridDim.x;
            }
            if (use > 0) mbar_wait(&empty_bar[s], (uint32_t)((use - 1) & 1ull));
            if (tile < 0) {
                if (lane == 0) {
                    item_tile[s] = -1;
                    mbar_arrive(&full_bar[s]);
                }
                break;
            }
            if (lane == 0) mbar_expect_tx(&full_bar[s], (uint32_t)tile_bytes);
            __syncwarp();
            const int chunk = a.tile_amps / a.copies;
            if (lane < a.copies)
                bulk_g2s(tiles + (size_t)s * a.tile_amps + lane * chunk, a.x + (u64)tile * a.tile_amps + (u64)lane * chunk,
                         (uint32_t)(chunk * sizeof(double2)), &full_bar[s]);
            if (lane == 0) {
                item_tile[s] = tile;
                mbar_arrive(&full_bar[s]);
            }
            __syncwarp();
        }
    } else {
        double acc = 0.0;
        const int per = a.tile_amps / (kConsumers * 32);  // amplitudes per thread
        for (u64 it = 0;; ++it) {
            const int s = (int)(it % a.stages);
            const u64 use = it / a.stages;
            mbar_wait(&full_bar[s], (uint32_t)(use & 1ull));
            const long long tile = item_tile[s];
            if (tile < 0) break;
            const double2* ts = tiles + (size_t)s * a.tile_amps;
            double2 v[8];
            for (int j0 = 0; j0 < per; j0 += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = (j0 + j < per) ? ts[t + (j0 + j) * kConsumers * 32] : make_double2(0.0, 0.0);
                for (int r = 1; r < a.lds_per_thread; ++r) {
                    const int tb = t ^ (1 << (r % 9));
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const double2 w = ts[tb + ((j0 + j) % per) * kConsumers * 32];
                        v[j].x = fma(0.5, w.x, v[j].x);
                        v[j].y = fma(0.5, w.y, v[j].y);
                    }
                }
                if (a.y) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j0 + j < per) a.y[(u64)tile * a.tile_amps + t + (j0 + j) * kConsumers * 32] = v[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc += v[j].x + v[j].y;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
        if (acc == 1.2345e300) a.out[0] = acc;
    }
}

static float time_it(cudaStream_t st, int reps, void (*launch)(void*), void* ctx) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch(ctx);
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(e0, st));
        launch(ctx);
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

struct Ctx {
    RingArgs a;
    int grid;
    size_t smem;
    const double2* x;
    double2* y;
    u64 n;
    double* out;
    int mode;
};
static void do_launch(void* p) {
    Ctx* c = (Ctx*)p;
    if (c->mode == 0) ldg_read<<<148 * 8, 256>>>(c->x, c->n, c->out);
    else if (c->mode == 1) ldg_copy<<<148 * 8, 256>>>(c->x, c->y, c->n);
    else {
        if (c->a.use_ticket) cudaMemsetAsync(c->a.ticket, 0, 4);
        ring_kernel<<<c->grid, kThreadsRing, c->smem>>>(c->a);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        printf("launch failed: %s\n", cudaGetErrorString(e));
        exit(1);
    }
}

int main(int argc, char** argv) {
    const int nq = argc > 1 ? atoi(argv[1]) : 25;
    const u64 n = 1ull << nq;
    double2 *x, *y;
    double* out;
    unsigned int* ticket;
    CK(cudaMalloc(&x, n * sizeof(double2)));
    CK(cudaMalloc(&y, n * sizeof(double2)));
    CK(cudaMalloc(&out, 64));
    CK(cudaMalloc(&ticket, 64));
    CK(cudaMemset(x, 0, n * sizeof(double2)));
    CK(cudaMemset(y, 0, n * sizeof(double2)));
    CK(cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const double gb = (double)n * 16 / 1e9;
    Ctx c{};
    c.x = x;
    c.y = y;
    c.n = n;
    c.out = out;
    c.mode = 0;
    float ms = time_it(0, 10, do_launch, &c);
    printf("ldg_read                                   %.4f ms  %.0f GB/s read\n", ms, gb / ms * 1e3);
    c.mode = 1;
    ms = time_it(0, 10, do_launch, &c);
    printf("ldg_copy                                   %.4f ms  %.0f GB/s read+write\n", ms, 2 * gb / ms * 1e3);
    c.mode = 2;
    const int tiles_kib[] = {64, 32, 16};
    for (int write = 0; write < 2; ++write)
        for (int ti = 0; ti < 3; ++ti)
            for (int ring_kib : {192, 96})
                for (int ticket_mode = 0; ticket_mode < 2; ++ticket_mode)
                    for (int lds : {1, 10}) {
                        const int tile_amps = tiles_kib[ti] * 1024 / 16;
                        const int stages = ring_kib / tiles_kib[ti];
                        if (stages < 2 || stages > 12) continue;
                        c.a = RingArgs{x, write ? y : nullptr, n / tile_amps, tile_amps, stages, 4, ticket_mode, ticket, lds, out};
                        c.grid = 148;
                        c.smem = (size_t)stages * tile_amps * 16;
                        ms = time_it(0, 6, do_launch, &c);
                        printf("ring tile %2d KiB x %2d stages ticket=%d lds=%2d write=%d   %.4f ms  %.0f GB/s %s\n", tiles_kib[ti], stages,
                               ticket_mode, lds, write, ms, (write ? 2 : 1) * gb / ms * 1e3, write ? "read+write" : "read");
                    }
    return 0;
}
