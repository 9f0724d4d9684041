/*
 * microbench.cu -- measures what bounds a hash-table build on this B200: random 32-byte sector
 * reads, L2 atomics (RED / ATOM with return / 128-bit CAS) as a function of the table size, so
 * that bench.py can report the graph build against a MEASURED random-access / atomic roofline
 * (SURVEY.md 8d "Atomic roof") instead of the streaming-copy figure only.
 *
 *   ./microbench [json-out]
 *
 * Every kernel touches one random 32-byte slot per operation (xorshift indices, full occupancy,
 * `ilp` independent operations in flight per thread).  Timed with CUDA events, best of 3.
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ u64 mix(u64 x) {
    x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32;
    return x;
}
__device__ __forceinline__ void ld_sector(const void *p, u64 &a, u64 &b, u64 &c, u64 &d) {
    asm volatile("ld.global.cg.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
}

enum Op { OP_LOAD = 0, OP_RED = 1, OP_ATOM = 2, OP_LOAD_RED = 3, OP_CAS128 = 4, OP_LOAD64B = 5 };

/* slots: 32-byte records; window: when non-zero, indices are random inside a window of `window`
 * slots that advances with the thread's progress (TLB-friendly, cache-unfriendly) */
template <int OP, int ILP>
__global__ void __launch_bounds__(256) k_rand(char *table, u64 n_slots, u64 ops_per_thread, u64 window, u64 *sink) {
    u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 nthreads = (u64)gridDim.x * blockDim.x;
    u64 acc = 0;
    for (u64 it = 0; it < ops_per_thread; it += ILP) {
        u64 idx[ILP];
#pragma unroll
        for (int u = 0; u < ILP; u++) {
            u64 h = mix((it + u) * nthreads + tid + 0x9E3779B97F4A7C15ull);
            if (window) {
                u64 base = (mix(it / 64 + 12345) % (n_slots / window)) * window; /* same window for ~64 iterations of all threads */
                idx[u] = base + __umul64hi(h, window);
            } else {
                idx[u] = __umul64hi(h, n_slots);
            }
        }
        if (OP == OP_LOAD || OP == OP_LOAD_RED || OP == OP_LOAD64B) {
            u64 a[ILP], b[ILP], c[ILP], d[ILP];
#pragma unroll
            for (int u = 0; u < ILP; u++) ld_sector(table + idx[u] * (OP == OP_LOAD64B ? 64 : 32), a[u], b[u], c[u], d[u]);
            if (OP == OP_LOAD64B) {
#pragma unroll
                for (int u = 0; u < ILP; u++) { u64 e, f, g, h; ld_sector(table + idx[u] * 64 + 32, e, f, g, h); acc += e ^ h; }
            }
#pragma unroll
            for (int u = 0; u < ILP; u++) {
                acc += a[u] ^ b[u] ^ c[u] ^ d[u];
                if (OP == OP_LOAD_RED) atomicAdd(reinterpret_cast<u32 *>(table + idx[u] * 32 + 16), 1u);
            }
        } else if (OP == OP_RED) {
#pragma unroll
            for (int u = 0; u < ILP; u++) atomicAdd(reinterpret_cast<u32 *>(table + idx[u] * 32 + 16), 1u);
        } else if (OP == OP_ATOM) {
            u32 r[ILP];
#pragma unroll
            for (int u = 0; u < ILP; u++) r[u] = atomicAdd(reinterpret_cast<u32 *>(table + idx[u] * 32 + 16), 1u);
#pragma unroll
            for (int u = 0; u < ILP; u++) acc += r[u];
        } else if (OP == OP_CAS128) {
#pragma unroll
            for (int u = 0; u < ILP; u++) {
                u64 lo, hi;
                asm volatile("{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 v, {%4, %5};\n\t"
                             "atom.global.relaxed.gpu.cas.b128 o, [%6], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
                             : "=l"(lo), "=l"(hi) : "l"(~0ull), "l"(~0ull), "l"(idx[u]), "l"(it), "l"(table + idx[u] * 32) : "memory");
                acc += lo ^ hi;
            }
        }
    }
    if (acc == 0x123456789ull) *sink = acc;
}

/* skew probe: every operation goes to one of `n_hot` slots (spread over the table) */
template <int OP>
__global__ void __launch_bounds__(256) k_hot(char *table, u64 n_slots, u64 n_hot, u64 ops_per_thread, u64 *sink) {
    u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 nthreads = (u64)gridDim.x * blockDim.x;
    u64 acc = 0;
    for (u64 it = 0; it < ops_per_thread; it++) {
        u64 h = mix(it * nthreads + tid + 0x9E3779B97F4A7C15ull);
        u64 idx = __umul64hi(mix((h % n_hot) + 77), n_slots);
        if (OP == OP_LOAD) { u64 a, b, c, d; ld_sector(table + idx * 32, a, b, c, d); acc += a ^ d; }
        else if (OP == OP_RED) atomicAdd(reinterpret_cast<u32 *>(table + idx * 32 + 16), 1u);
        else if (OP == OP_ATOM) acc += atomicAdd(reinterpret_cast<u32 *>(table + idx * 32 + 16), 1u);
    }
    if (acc == 0x123456789ull) *sink = acc;
}
template <int OP>
double run_hot(char *table, u64 n_slots, u64 n_hot, u64 *sink, int sms) {
    const u64 total_ops = 1ull << 24;
    int grid = sms * 8;
    u64 per_thread = total_ops / ((u64)grid * 256);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        k_hot<OP><<<grid, 256>>>(table, n_slots, n_hot, per_thread, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return (double)per_thread * grid * 256 / (best * 1e-3) / 1e9;
}

template <int OP, int ILP>
double run(char *table, u64 n_slots, u64 window, u64 *sink, int sms, int blocks_per_sm) {
    const u64 total_ops = 1ull << 28;
    int grid = sms * blocks_per_sm;
    u64 per_thread = total_ops / ((u64)grid * 256);
    per_thread -= per_thread % ILP;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        k_rand<OP, ILP><<<grid, 256>>>(table, n_slots, per_thread, window, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return (double)per_thread * grid * 256 / (best * 1e-3) / 1e9; /* G ops/s */
}

int main(int argc, char **argv) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const size_t max_bytes = 8ull << 30;
    char *table; u64 *sink;
    CK(cudaMalloc(&table, max_bytes));
    CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(table, 0xFF, max_bytes));
    FILE *out = argc > 1 ? fopen(argv[1], "w") : stdout;
    fprintf(out, "{\"gpu\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"unit\": \"G ops/s\", \"rows\": [\n", prop.name, sms, prop.l2CacheSize);
    const bool quick = argc > 2;
    const size_t sizes[] = { 32ull << 20, 96ull << 20, 256ull << 20, 512ull << 20, 1ull << 30, 2ull << 30, 8ull << 30 };
    bool first = true;
    for (size_t bytes : sizes) {
        if (quick && bytes != (32ull << 20)) continue;
        u64 n = bytes / 32;
        double load1 = run<OP_LOAD, 1>(table, n, 0, sink, sms, 8);
        double load4 = run<OP_LOAD, 4>(table, n, 0, sink, sms, 8);
        double load8 = run<OP_LOAD, 8>(table, n, 0, sink, sms, 6);
        double red1 = run<OP_RED, 1>(table, n, 0, sink, sms, 8);
        double red4 = run<OP_RED, 4>(table, n, 0, sink, sms, 8);
        double atom1 = run<OP_ATOM, 1>(table, n, 0, sink, sms, 8);
        double atom4 = run<OP_ATOM, 4>(table, n, 0, sink, sms, 8);
        double lr1 = run<OP_LOAD_RED, 1>(table, n, 0, sink, sms, 8);
        double lr4 = run<OP_LOAD_RED, 4>(table, n, 0, sink, sms, 8);
        double cas4 = run<OP_CAS128, 4>(table, n, 0, sink, sms, 8);
        double l64 = run<OP_LOAD64B, 4>(table, n / 2, 0, sink, sms, 8);
        /* TLB probe: same footprint, but all threads stay inside a 64 MB window at a time */
        double loadw = bytes > (64ull << 20) ? run<OP_LOAD, 4>(table, n, (64ull << 20) / 32, sink, sms, 8) : load4;
        double lrw = bytes > (64ull << 20) ? run<OP_LOAD_RED, 4>(table, n, (64ull << 20) / 32, sink, sms, 8) : lr4;
        fprintf(out, "%s{\"table_mb\": %zu, \"load_ilp1\": %.2f, \"load_ilp4\": %.2f, \"load_ilp8\": %.2f, \"red_ilp1\": %.2f, \"red_ilp4\": %.2f, "
                     "\"atom_ilp1\": %.2f, \"atom_ilp4\": %.2f, \"load_red_ilp1\": %.2f, \"load_red_ilp4\": %.2f, \"cas128_ilp4\": %.2f, "
                     "\"load64B_ilp4\": %.2f, \"load_ilp4_window64MB\": %.2f, \"load_red_ilp4_window64MB\": %.2f}",
                first ? "" : ",\n", bytes >> 20, load1, load4, load8, red1, red4, atom1, atom4, lr1, lr4, cas4, l64, loadw, lrw);
        fflush(out);
        first = false;
        CK(cudaMemset(table, 0xFF, max_bytes));
    }
    fprintf(out, "\n],\n\"hot_addresses\": [\n");
    {
        const u64 n = (32ull << 20) / 32;
        const u64 hots[] = { 1, 4, 16, 64, 256, 1024, 4096, 65536 };
        bool f2 = true;
        for (u64 nh : hots) {
            double l = run_hot<OP_LOAD>(table, n, nh, sink, sms);
            double r = run_hot<OP_RED>(table, n, nh, sink, sms);
            double a = run_hot<OP_ATOM>(table, n, nh, sink, sms);
            fprintf(out, "%s{\"n_hot\": %llu, \"load\": %.3f, \"red\": %.3f, \"atom\": %.3f}", f2 ? "" : ",\n", nh, l, r, a);
            fflush(out);
            f2 = false;
        }
    }
    fprintf(out, "\n]}\n");
    if (out != stdout) fclose(out);
    return 0;
}
