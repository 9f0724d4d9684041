// Experiment (round 1): how fast can only the EVEN records of a page-locked text buffer reach the device?
//   (a) contiguous cudaMemcpyAsync of everything (what staging does today)
//   (b) cudaMemcpy2DAsync: width = record, source pitch = 2 records
//   (c) zero-copy: a kernel reads the even records straight from mapped host memory (warp per record)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o h2d_strided h2d_strided.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_pull(const unsigned char *host, unsigned char *dev, size_t n_pairs, int rec_len) {
    const unsigned lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t i = warp; i < n_pairs; i += n_warps) {
        const unsigned char *src = host + i * 2 * rec_len;
        unsigned char *dst = dev + i * rec_len;
        for (int j = lane; j < rec_len; j += 32) dst[j] = src[j];
    }
}
// same, 16-byte loads where the alignment allows (head/tail bytes singly)
__global__ void k_pull16(const unsigned char *host, unsigned char *dev, size_t n_pairs, int rec_len) {
    const unsigned lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t i = warp; i < n_pairs; i += n_warps) {
        const unsigned char *src = host + i * 2 * rec_len;
        unsigned char *dst = dev + i * rec_len;
        const size_t a = (size_t)src & 15;
        const int head = a ? (int)(16 - a) : 0;
        const int body = (rec_len - head) / 16;
        if ((int)lane < head) dst[lane] = src[lane];
        if ((int)lane < body) {
            uint4 v = *reinterpret_cast<const uint4 *>(src + head + lane * 16);
            unsigned char *d = dst + head + lane * 16;
            unsigned w[4] = { v.x, v.y, v.z, v.w };
            for (int b = 0; b < 16; b++) d[b] = (unsigned char)(w[b >> 2] >> ((b & 3) * 8));
        }
        const int tail0 = head + body * 16;
        if (tail0 + (int)lane < rec_len) dst[tail0 + lane] = src[tail0 + lane];
    }
}

int main() {
    const int L = 50, rec_len = 2 * L + 1;
    const size_t n_rec = 20000000, n_pairs = n_rec / 2, bytes = n_rec * rec_len;
    unsigned char *h, *d;
    CK(cudaHostAlloc(&h, bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    memset(h, 'A', bytes);
    CK(cudaMalloc(&d, bytes));
    cudaStream_t s[4];
    for (auto &x : s) CK(cudaStreamCreate(&x));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0, s[0]));
        CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s[0]));
        CK(cudaEventRecord(e1, s[0])); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("contiguous   %7.2f ms  %6.1f GB/s of text\n", ms, bytes / ms / 1e6);
        CK(cudaEventRecord(e0, s[0]));
        CK(cudaMemcpy2DAsync(d, rec_len, h, 2 * rec_len, rec_len, n_pairs, cudaMemcpyHostToDevice, s[0]));
        CK(cudaEventRecord(e1, s[0])); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("memcpy2D     %7.2f ms  %6.1f GB/s of even records (%5.1f GB/s text-equivalent)\n", ms, bytes / 2 / ms / 1e6, bytes / ms / 1e6);
        for (int grid : { 148 * 4, 148 * 16 }) {
            CK(cudaEventRecord(e0, s[0]));
            k_pull<<<grid, 256, 0, s[0]>>>(h, d, n_pairs, rec_len);
            CK(cudaEventRecord(e1, s[0])); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("zero-copy    %7.2f ms  %6.1f GB/s of even records (grid %d)\n", ms, bytes / 2 / ms / 1e6, grid);
            CK(cudaEventRecord(e0, s[0]));
            k_pull16<<<grid, 256, 0, s[0]>>>(h, d, n_pairs, rec_len);
            CK(cudaEventRecord(e1, s[0])); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("zero-copy16  %7.2f ms  %6.1f GB/s of even records (grid %d)\n", ms, bytes / 2 / ms / 1e6, grid);
        }
        // the whole text pulled by the kernel, for reference
        CK(cudaEventRecord(e0, s[0]));
        k_pull<<<148 * 16, 256, 0, s[0]>>>(h, d, n_rec, rec_len / 2 + 0);   // (half-length records: every byte of the first half)
        CK(cudaEventRecord(e1, s[0])); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("zero-copy 50B runs %7.2f ms\n", ms);
    }
    CK(cudaGetLastError());
    return 0;
}
