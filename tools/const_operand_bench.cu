// Microbenchmark for DESIGN.md section 9 item 1: how fast is the tensor-build phase of the assembly kernel when the
// element-independent tensor values are (a) constant-bank immediates, (b) constant-bank values behind a warp-uniform
// run-time index (LDC), against (c) today's lane-dependent shared-memory reads?  Every variant forms, per element,
// 320 matrix entries x 14 tensors = 4 480 DFMAs (p = 4 needs 4 500) and stores the entries to shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/const_operand_bench tools/const_operand_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NT = 14, NE = 320;            // tensors, entries per element
__constant__ double c_tab[NT * NE];

// (a)/(b): 512 threads = 16 warps build 32 elements per iteration: warp w owns entries [20 w, 20 w + 20), lane = element
template <bool IMM>
__global__ void __launch_bounds__(512, 1) build_const(int iters, const double* __restrict__ coef_g, double* __restrict__ out, int base_rt) {
    extern __shared__ double sm[];          // [32 elements][NE + 1]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double acc_keep = 0.0;
    for (int it = blockIdx.x; it < iters; it += gridDim.x) {
        double c[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) c[t] = coef_g[(size_t)(it * 32 + lane) * NT + t];
        const int base = IMM ? 0 : base_rt + w * 20;       // IMM: every warp the same compile-time entries (throughput only)
#pragma unroll
        for (int q = 0; q < 20; q += 2) {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                a0 += c[t] * c_tab[t * NE + base + q];
                a1 += c[t] * c_tab[t * NE + base + q + 1];
            }
            *reinterpret_cast<double2*>(&sm[lane * (NE + 2) + w * 20 + q]) = make_double2(a0, a1);
            acc_keep += a0 + a1;
        }
        __syncthreads();
    }
    if (acc_keep == 12345.678) out[threadIdx.x] = sm[threadIdx.x];
}

// (c): today's layout: 128 threads, 16 lanes per element, lane = row, 20 entries per lane, tensor rows in shared memory
__global__ void __launch_bounds__(128, 4) build_smem(int iters, const double* __restrict__ coef_g, const double* __restrict__ tab_g,
                                                      double* __restrict__ out) {
    extern __shared__ double sm[];          // [NT][16 rows][21]
    for (int i = threadIdx.x; i < NT * 16 * 21; i += 128) sm[i] = tab_g[i % (NT * NE)];
    __syncthreads();
    const int row = threadIdx.x & 15, grp = threadIdx.x >> 4;
    double acc_keep = 0.0;
    for (int it = blockIdx.x; it < iters; it += gridDim.x) {       // 8 elements per iteration
        double c[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) c[t] = coef_g[(size_t)(it * 8 + grp) * NT + t];
#pragma unroll 2
        for (int q = 0; q < 20; ++q) {
            double a0 = 0.0;
#pragma unroll
            for (int t = 0; t < NT; ++t) a0 += c[t] * sm[(t * 16 + row) * 21 + q];
            acc_keep += a0;
        }
    }
    if (acc_keep == 12345.678) out[threadIdx.x] = acc_keep;
}

int main() {
    const int n_el = 1 << 20;
    double *coef, *tab, *out;
    cudaMalloc(&coef, sizeof(double) * (size_t)n_el * NT);
    cudaMalloc(&tab, sizeof(double) * NT * NE);
    cudaMalloc(&out, sizeof(double) * 1024);
    cudaMemset(coef, 0, sizeof(double) * (size_t)n_el * NT);
    double h[NT * NE];
    for (int i = 0; i < NT * NE; ++i) h[i] = 1.0 / (1 + i);
    cudaMemcpy(tab, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(c_tab, h, sizeof(h));
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const size_t sm_const = sizeof(double) * 32 * (NE + 2), sm_tab = sizeof(double) * NT * 16 * 21;
    cudaFuncSetAttribute(build_const<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_const);
    cudaFuncSetAttribute(build_const<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_const);
    cudaFuncSetAttribute(build_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_tab);
    for (int variant = 0; variant < 3; ++variant) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (variant == 0) build_const<true><<<sms, 512, sm_const>>>(n_el / 32, coef, out, 0);
            if (variant == 1) build_const<false><<<sms, 512, sm_const>>>(n_el / 32, coef, out, 0);
            if (variant == 2) build_smem<<<sms * 4, 128, sm_tab>>>(n_el / 8, coef, tab, out);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        const char* names[] = {"constant bank, immediate offsets", "constant bank, warp-uniform run-time index (LDC)", "shared memory, lane = row (today)"};
        printf("{\"variant\": \"%s\", \"ms_per_1M_elements\": %.3f, \"tflops\": %.2f, \"err\": \"%s\"}\n", names[variant], best,
               2.0 * NT * NE * n_el / (best * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
