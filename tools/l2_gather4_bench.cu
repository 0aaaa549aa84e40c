// Micro-benchmark: random ROW gathers from an L2-resident table into shared memory with the Blackwell TMA
// tile::gather4 mode (one instruction = four rows of a 2-D tensor, `UTMALDG.2D.GATHER4`), against the 1-D bulk copy
// (`UBLKCP`, one instruction per row) measured in round 1 (l2_bulk_gather_bench.cu: 43 G rows/s at best).
// Question it answers: is the TMA row rate bound per INSTRUCTION (then gather4 is up to 4x faster and the den kernels
// can stage their gathered rows in shared memory instead of registers) or per ROW (then nothing changes)?
// Each warp owns a ring of STAGES stages of R rows; lane q of the first R/4 lanes issues one gather4 per stage; the warp
// then reads the rows with LDS and does the recursion's two FMAs per lane and row.  Results are verified (row r of the
// table holds the value r).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2_gather4_bench tools/l2_gather4_bench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *tm, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

// ISSUERS: 0 = lane q issues quad q of the stage (R/4 lanes active in one instruction); 1 = lane 0 issues all R/4 quads
template <int ROWB, int R, int STAGES, int ISSUERS>
__global__ void g4_gather(const __grid_constant__ CUtensorMap tm, const unsigned *idx, int per_warp, double *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    unsigned char *buf = smem + (size_t)warp * STAGES * R * ROWB;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)nw * STAGES * R * ROWB) + warp * STAGES;
    if (lane == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(bars + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const unsigned *my = idx + ((size_t)blockIdx.x * nw + warp) * per_warp;
    const int n = per_warp / R;
    auto issue = [&](int k) {
        const int s = k % STAGES;
        const uint32_t bar = smem_u32(bars + s);
        if (lane == 0) mbar_expect_tx(bar, R * ROWB);
        __syncwarp();
        if (ISSUERS == 0) {
            if (lane < R / 4) {
                const uint4 q = __ldg(reinterpret_cast<const uint4 *>(my + k * R) + lane);
                tma_gather4(smem_u32(buf + (size_t)(s * R + 4 * lane) * ROWB), &tm, 0, q.x, q.y, q.z, q.w, bar);
            }
        } else if (lane == 0) {
#pragma unroll
            for (int j = 0; j < R / 4; ++j) {
                const uint4 q = __ldg(reinterpret_cast<const uint4 *>(my + k * R) + j);
                tma_gather4(smem_u32(buf + (size_t)(s * R + 4 * j) * ROWB), &tm, 0, q.x, q.y, q.z, q.w, bar);
            }
        }
    };
    for (int k = 0; k < STAGES && k < n; ++k) issue(k);
    float a0 = 0.f, a1 = 0.f;
    double chk = 0.0;
    for (int k = 0; k < n; ++k) {
        const int s = k % STAGES;
        mbar_wait(smem_u32(bars + s), (k / STAGES) & 1);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (ROWB >= 256) {
                const float2 v = *reinterpret_cast<const float2 *>(buf + (size_t)(s * R + r) * ROWB + lane * 8);
                a0 = fmaf(v.x, 1.0f, a0); a1 = fmaf(v.y, 1.0f, a1);
            } else {
                const float v = *reinterpret_cast<const float *>(buf + (size_t)(s * R + r) * ROWB + lane * 4);
                a0 = fmaf(v, 1.0f, a0);
            }
        }
        if ((k & 63) == 63) { chk += (double)a0 + (double)a1; a0 = 0.f; a1 = 0.f; }
        __syncwarp();
        if (k + STAGES < n) issue(k + STAGES);
    }
    chk += (double)a0 + (double)a1;
    if (lane == 0) out[(size_t)blockIdx.x * nw + warp] = chk;   // = (ROWB>=256 ? 2 : 1) * sum of the row indices
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Bench {
    int sms, rows, rowb;
    float *tab; double *out; unsigned *idx;
    std::vector<unsigned> h;
    static constexpr int per_warp = 4096;
    EncodeTiled encode;
    void setup(int rows_, int rowb_, int max_warps) {
        rows = rows_; rowb = rowb_;
        std::vector<float> t((size_t)rows * rowb / 4);
        for (int r = 0; r < rows; ++r) for (int c = 0; c < rowb / 4; ++c) t[(size_t)r * (rowb / 4) + c] = (float)r;
        cudaMalloc(&tab, t.size() * 4); cudaMemcpy(tab, t.data(), t.size() * 4, cudaMemcpyHostToDevice);
        cudaMalloc(&out, (size_t)max_warps * 8);
        h.resize((size_t)max_warps * per_warp);
        unsigned s = 12345;
        for (auto &x : h) { s = s * 1664525u + 1013904223u; x = (s >> 8) % rows; }
        cudaMalloc(&idx, h.size() * 4); cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        encode = (EncodeTiled)fn;
    }
    bool make_map(CUtensorMap *tm, int box_rows) {
        cuuint64_t dims[2] = {(cuuint64_t)(rowb / 4), (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)rowb};
        cuuint32_t box[2] = {(cuuint32_t)(rowb / 4), (cuuint32_t)box_rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, tab, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled(box rows %d) failed: %d\n", box_rows, (int)r); return false; }
        return true;
    }
    void teardown() { cudaFree(tab); cudaFree(out); cudaFree(idx); }
};

template <int ROWB, int R, int STAGES, int ISSUERS>
void run(Bench &b, int threads, int box_rows) {
    const int nw = threads / 32;
    const size_t smem = (size_t)nw * STAGES * R * ROWB + (size_t)nw * STAGES * 8;
    auto kern = g4_gather<ROWB, R, STAGES, ISSUERS>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        printf("gather4 rowB=%d R=%d stages=%d threads=%d: smem %zu too large\n", ROWB, R, STAGES, threads, smem);
        cudaGetLastError();
        return;
    }
    CUtensorMap tm;
    if (!b.make_map(&tm, box_rows)) return;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaMemset(b.out, 0, (size_t)b.sms * nw * 8);
    kern<<<b.sms, threads, smem>>>(tm, b.idx, b.per_warp, b.out);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("gather4 rowB=%d R=%d stages=%d threads=%d box_rows=%d: %s\n", ROWB, R, STAGES, threads, box_rows, cudaGetErrorString(err)); exit(1); }
    // verify
    std::vector<double> got((size_t)b.sms * nw);
    cudaMemcpy(got.data(), b.out, got.size() * 8, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (size_t w = 0; w < got.size(); ++w) {
        double want = 0.0;
        for (int i = 0; i < b.per_warp; ++i) want += (double)b.h[w * b.per_warp + i];
        want *= (ROWB >= 256 ? 2.0 : 1.0);
        if (got[w] != want) ++bad;
    }
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) kern<<<b.sms, threads, smem>>>(tm, b.idx, b.per_warp, b.out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double nrows = (double)b.sms * nw * b.per_warp;
    printf("gather4 rowB=%d R=%d stages=%d threads=%d issue=%s box_rows=%d inflight=%zuKB : %.3f ms  %.1f GB/s  %.2f Grows/s  wrong_warps=%zu/%zu (%s)\n",
           ROWB, R, STAGES, threads, ISSUERS ? "lane0" : "lanes", box_rows, (size_t)nw * STAGES * R * ROWB >> 10, ms,
           nrows * ROWB / ms * 1e-6, nrows / ms * 1e-6, bad, got.size(), cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
}

int main(int argc, char **argv) {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    Bench b; b.sms = p.multiProcessorCount;
    printf("%s SMs=%d\n", p.name, b.sms);
    const int box_rows = argc > 1 ? atoi(argv[1]) : 1;
    b.setup(60000, 256, b.sms * 32);
    run<256, 16, 2, 0>(b, 512, box_rows);  run<256, 16, 3, 0>(b, 512, box_rows);  run<256, 16, 2, 1>(b, 512, box_rows);
    run<256, 32, 2, 0>(b, 256, box_rows);  run<256, 32, 3, 0>(b, 256, box_rows);  run<256, 32, 2, 1>(b, 256, box_rows);
    run<256, 32, 4, 0>(b, 128, box_rows);  run<256, 64, 2, 0>(b, 128, box_rows);  run<256, 64, 2, 1>(b, 128, box_rows);
    run<256, 8, 2, 0>(b, 1024, box_rows);  run<256, 8, 3, 0>(b, 1024, box_rows);  run<256, 16, 1, 0>(b, 1024, box_rows);
    run<256, 16, 3, 1>(b, 512, box_rows);  run<256, 8, 6, 1>(b, 512, box_rows);
    b.teardown();
    b.setup(60000, 128, b.sms * 32);
    run<128, 16, 4, 0>(b, 512, box_rows);  run<128, 32, 3, 0>(b, 512, box_rows);  run<128, 32, 3, 1>(b, 512, box_rows);
    run<128, 16, 3, 0>(b, 1024, box_rows); run<128, 64, 3, 0>(b, 256, box_rows);
    b.teardown();
    b.setup(60000, 512, b.sms * 32);
    run<512, 16, 1, 0>(b, 512, box_rows);  run<512, 8, 3, 0>(b, 512, box_rows);  run<512, 16, 2, 0>(b, 256, box_rows);
    b.teardown();
    return 0;
}
