// blur_ab.cu — standalone A/B microbenchmark for the Gaussian blur of the feature-extraction stage (the kernel
// is at ~6 % of its HBM roofline).  NOT product code: variants are tried here first, the winner moves to csrc/sift.cu.
//
//   v0  the product kernel as of round 1 (tile + halo -> smem, x pass -> smem, y pass -> HBM; 8 outputs per thread)
//   v1  same arithmetic, cheaper tile load: 2-D thread mapping (no integer division), row pointer hoisted, interior tiles
//       skip the reflect-101 index arithmetic                                   -> must equal v0 bit for bit
//   v2  v1 + radius as a template parameter: tap loops fully unrolled, weights read as constant-bank operands
//                                                                               -> must equal v0 bit for bit
//   v3  v2 with fused multiply-add (what an FMA build of OpenCV does)          -> differs in the last bits, max |diff| printed
//
// Build + run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --fmad=false -o tools/blur_ab tools/blur_ab.cu && tools/blur_ab
// (v3 uses explicit fmaf, so --fmad=false keeps v0-v2 on separate multiply / add as in the product build.)
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int kMaxR = 32;
constexpr int TW = 64, TH = 32, MS = TW + 1;
struct Weights { int radius; float w[2 * kMaxR + 1]; };

__host__ __device__ inline int in_stride(int R) { return (TW + 2 * R) | 1; }
__device__ __forceinline__ int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}

// ------------------------------------------------------------------------------------------------ shared pieces
template <bool FMA>
__device__ __forceinline__ void eight_outputs(const float* __restrict__ p, int stride, int taps, const Weights& k, float acc[8]) {
    float win[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) { acc[o] = 0.f; win[o] = p[o * stride]; }
    for (int t8 = 0; t8 < taps; t8 += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (t8 + u < taps) {
                const float wt = k.w[t8 + u];
#pragma unroll
                for (int o = 0; o < 8; ++o) acc[o] = FMA ? fmaf(wt, win[(o + u) & 7], acc[o]) : acc[o] + wt * win[(o + u) & 7];
                win[u] = p[(t8 + u + 8) * stride];
            }
        }
    }
}
template <int R, bool FMA>
__device__ __forceinline__ void eight_outputs_fixed(const float* __restrict__ p, int stride, const Weights& k, float acc[8]) {
    constexpr int taps = 2 * R + 1;
    float win[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) { acc[o] = 0.f; win[o] = p[o * stride]; }
#pragma unroll
    for (int t = 0; t < taps; ++t) {
        const float wt = k.w[t];
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[o] = FMA ? fmaf(wt, win[(o + t) & 7], acc[o]) : acc[o] + wt * win[(o + t) & 7];
        win[t & 7] = p[(t + 8) * stride];
    }
}

__device__ __forceinline__ void load_tile_v0(const float* __restrict__ src, int w, int h, int x0, int y0, int R, float* in) {
    const int IW = TW + 2 * R, IH = TH + 2 * R, IS = in_stride(R);
    for (int idx = threadIdx.x; idx < IH * IW; idx += 256) {
        const int iy = idx / IW, ix = idx - iy * IW;
        in[iy * IS + ix] = src[static_cast<size_t>(reflect101(y0 - R + iy, h)) * w + reflect101(x0 - R + ix, w)];
    }
}
__device__ __forceinline__ void load_tile_v1(const float* __restrict__ src, int w, int h, int x0, int y0, int R, float* in) {
    const int IW = TW + 2 * R, IH = TH + 2 * R, IS = in_stride(R);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const bool interior = x0 - R >= 0 && x0 + TW + R <= w && y0 - R >= 0 && y0 + TH + R <= h;
    if (interior) {
        const float* base = src + static_cast<size_t>(y0 - R) * w + (x0 - R);
        for (int iy = ty; iy < IH; iy += 8) {
            const float* row = base + static_cast<size_t>(iy) * w;
            float* dst = in + iy * IS;
            for (int ix = tx; ix < IW; ix += 32) dst[ix] = row[ix];
        }
    } else {
        for (int iy = ty; iy < IH; iy += 8) {
            const float* row = src + static_cast<size_t>(reflect101(y0 - R + iy, h)) * w;
            float* dst = in + iy * IS;
            for (int ix = tx; ix < IW; ix += 32) dst[ix] = row[reflect101(x0 - R + ix, w)];
        }
    }
}

// ------------------------------------------------------------------------------------------------ variants
template <int VARIANT>
__global__ void __launch_bounds__(256) blur_generic(const float* __restrict__ src, float* __restrict__ dst, int w, int h, const Weights k) {
    extern __shared__ float sm[];
    const int R = k.radius, IH = TH + 2 * R, IS = in_stride(R);
    float* in = sm;
    float* mid = sm + IH * IS + 8;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, tid = threadIdx.x;
    if (VARIANT == 0) load_tile_v0(src, w, h, x0, y0, R, in); else load_tile_v1(src, w, h, x0, y0, R, in);
    __syncthreads();
    const int taps = 2 * R + 1;
    for (int unit = tid; unit < IH * (TW / 8); unit += 256) {
        const int iy = unit % IH, ix0 = (unit / IH) * 8;
        float acc[8];
        eight_outputs<false>(in + iy * IS + ix0, 1, taps, k, acc);
#pragma unroll
        for (int o = 0; o < 8; ++o) mid[iy * MS + ix0 + o] = acc[o];
    }
    __syncthreads();
    const int ox = tid % TW, oy0 = (tid / TW) * 8;
    float acc[8];
    eight_outputs<false>(mid + oy0 * MS + ox, MS, taps, k, acc);
    if (x0 + ox < w)
#pragma unroll
        for (int o = 0; o < 8; ++o)
            if (y0 + oy0 + o < h) dst[static_cast<size_t>(y0 + oy0 + o) * w + x0 + ox] = acc[o];
}

template <int R, bool FMA>
__global__ void __launch_bounds__(256) blur_fixed(const float* __restrict__ src, float* __restrict__ dst, int w, int h, const Weights k) {
    extern __shared__ float sm[];
    constexpr int IH = TH + 2 * R;
    const int IS = in_stride(R);
    float* in = sm;
    float* mid = sm + IH * IS + 8;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, tid = threadIdx.x;
    load_tile_v1(src, w, h, x0, y0, R, in);
    __syncthreads();
    for (int unit = tid; unit < IH * (TW / 8); unit += 256) {
        const int iy = unit % IH, ix0 = (unit / IH) * 8;
        float acc[8];
        eight_outputs_fixed<R, FMA>(in + iy * IS + ix0, 1, k, acc);
#pragma unroll
        for (int o = 0; o < 8; ++o) mid[iy * MS + ix0 + o] = acc[o];
    }
    __syncthreads();
    const int ox = tid % TW, oy0 = (tid / TW) * 8;
    float acc[8];
    eight_outputs_fixed<R, FMA>(mid + oy0 * MS + ox, MS, k, acc);
    if (x0 + ox < w)
#pragma unroll
        for (int o = 0; o < 8; ++o)
            if (y0 + oy0 + o < h) dst[static_cast<size_t>(y0 + oy0 + o) * w + x0 + ox] = acc[o];
}

// ------------------------------------------------------------------------------------------------ driver
static Weights make_weights(double sigma) {
    Weights k{};
    const int ksize = static_cast<int>(std::lrint(sigma * 8 + 1)) | 1;
    k.radius = (ksize - 1) / 2;
    std::vector<double> v(ksize);
    double sum = 0;
    for (int i = 0; i < ksize; ++i) { const double x = i - k.radius; v[i] = std::exp(-(x * x) / (2 * sigma * sigma)); sum += v[i]; }
    for (int i = 0; i < ksize; ++i) k.w[i] = static_cast<float>(v[i] / sum);
    return k;
}
static size_t smem_bytes(int R) { const size_t ih = TH + 2 * R; return (ih * in_stride(R) + 8 + (ih + 8) * MS) * sizeof(float); }

template <class F>
static float time_us(F launch, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return 1e3f * ms / reps;
}

template <int R>
static void run_radius(const float* d_src, float* d_out, float* d_ref, int w, int h, double sigma, const std::vector<float>& host_src) {
    const Weights k = make_weights(sigma);
    if (k.radius != R) { std::printf("sigma %.4f gives radius %d, expected %d\n", sigma, k.radius, R); return; }
    const dim3 grid((w + TW - 1) / TW, (h + TH - 1) / TH);
    const size_t smem = smem_bytes(R), n = static_cast<size_t>(w) * h;
    cudaFuncSetAttribute(blur_generic<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(blur_generic<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(blur_fixed<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(blur_fixed<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    std::vector<float> ref(n), got(n);
    blur_generic<0><<<grid, 256, smem>>>(d_src, d_ref, w, h, k);
    cudaMemcpy(ref.data(), d_ref, n * 4, cudaMemcpyDeviceToHost);
    auto check = [&](const char* name, bool exact) {
        cudaMemcpy(got.data(), d_out, n * 4, cudaMemcpyDeviceToHost);
        double md = 0;
        size_t bad = 0;
        for (size_t i = 0; i < n; ++i) { const double d = std::fabs(static_cast<double>(got[i]) - ref[i]); md = d > md ? d : md; bad += got[i] != ref[i]; }
        std::printf("    %-28s max |diff| vs v0 %.3g, differing elements %zu%s\n", name, md, bad, exact && bad ? "   <-- NOT bit-identical" : "");
    };
    const double gb = 8.0 * n / 1e9;
    const float t0 = time_us([&] { blur_generic<0><<<grid, 256, smem>>>(d_src, d_out, w, h, k); }, 20);
    std::printf("  R = %2d (%d taps), %d x %d: v0 %.1f us (%.0f GB/s)\n", R, 2 * R + 1, w, h, t0, gb / (t0 * 1e-6));
    const float t1 = time_us([&] { blur_generic<1><<<grid, 256, smem>>>(d_src, d_out, w, h, k); }, 20);
    std::printf("    v1 cheap tile load          %.1f us (%.0f GB/s)\n", t1, gb / (t1 * 1e-6));
    check("v1", true);
    const float t2 = time_us([&] { blur_fixed<R, false><<<grid, 256, smem>>>(d_src, d_out, w, h, k); }, 20);
    std::printf("    v2 + unrolled taps          %.1f us (%.0f GB/s)\n", t2, gb / (t2 * 1e-6));
    check("v2", true);
    const float t3 = time_us([&] { blur_fixed<R, true><<<grid, 256, smem>>>(d_src, d_out, w, h, k); }, 20);
    std::printf("    v3 + fused multiply-add     %.1f us (%.0f GB/s)\n", t3, gb / (t3 * 1e-6));
    check("v3", false);
    (void)host_src;
}

int main(int argc, char** argv) {
    const int w = argc > 2 ? std::atoi(argv[1]) : 3200, h = argc > 2 ? std::atoi(argv[2]) : 2400;
    const size_t n = static_cast<size_t>(w) * h;
    std::vector<float> src(n);
    unsigned s = 12345;
    for (size_t i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; src[i] = static_cast<float>((s >> 8) & 0xffff) * (255.f / 65535.f); }
    float *d_src, *d_out, *d_ref;
    cudaMalloc(&d_src, n * 4); cudaMalloc(&d_out, n * 4); cudaMalloc(&d_ref, n * 4);
    cudaMemcpy(d_src, src.data(), n * 4, cudaMemcpyHostToDevice);
    std::printf("Gaussian blur variants, %d x %d fp32 (algorithmic traffic 8 B / pixel = %.1f MB)\n", w, h, 8.0 * n / 1e6);
    // the radii of cv::SIFT's default pyramid (sigma 1.6, 3 layers): 11, 11, 13, 17, 21, 27 taps
    run_radius<5>(d_src, d_out, d_ref, w, h, 1.2262734984654078, src);
    run_radius<6>(d_src, d_out, d_ref, w, h, 1.5450077936447955, src);
    run_radius<8>(d_src, d_out, d_ref, w, h, 1.9465878414647133, src);
    run_radius<10>(d_src, d_out, d_ref, w, h, 2.4525469969308156, src);
    run_radius<13>(d_src, d_out, d_ref, w, h, 3.0900155872895910, src);
    const cudaError_t e = cudaDeviceSynchronize();
    std::printf("%s\n", e == cudaSuccess ? "done" : cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 1;
}
