// tma_probe.cu — stand-alone probe of TMA tensor-map geometries (debug tool, not product).
// usage: tma_probe <case>   (each case in its own process: CUDA errors are sticky)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap map, float* out, int nfloats, int c0, int c1, int c2, int c3) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + ((nfloats * 4 + 127) & ~127));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(nfloats * 4) : "memory");
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(smem_u32(smem)), "l"(&map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(smem_u32(smem)), "l"(&map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
        if (RANK == 4)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(smem_u32(smem)), "l"(&map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n"
        ::"r"(smem_u32(bar)), "r"(0) : "memory");
    const float* s = reinterpret_cast<const float*>(smem);
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = s[i];
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

int main(int argc, char** argv) {
    int cs = argc > 1 ? atoi(argv[1]) : 0;
    const int W = 256, Wp = 256, H = 144, Hp = 144;
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)p;
    std::vector<float> h((size_t)Wp * Hp * 4);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
    float *d, *out;
    CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&out, 1 << 20));
    CUtensorMap m; CUresult r; int rank = 0, nfl = 0; int S = 1;
    const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    int c[4] = {0, 0, 0, 0};
    if (cs == 0) {  // 2-D float plane (x, y), box 132 x 12, interior
        rank = 2; cuuint64_t dims[2] = {W, H}, str[1] = {(cuuint64_t)Wp * 4}; cuuint32_t box[2] = {132, 12};
        r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        nfl = 132 * 12; c[0] = 4; c[1] = 4;
    } else if (cs == 1) {  // same with negative start
        rank = 2; cuuint64_t dims[2] = {W, H}, str[1] = {(cuuint64_t)Wp * 4}; cuuint32_t box[2] = {132, 12};
        r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        nfl = 132 * 12; c[0] = -2; c[1] = -2;
    } else if (cs == 2 || cs == 3) {  // 3-D (x, phase, k) float plane; S = 1 or 4
        S = cs == 2 ? 1 : 4; rank = 3;
        cuuint64_t dims[3] = {W, (cuuint64_t)S, (cuuint64_t)(Hp / S)}, str[2] = {(cuuint64_t)Wp * 4, (cuuint64_t)Wp * 4 * S};
        cuuint32_t box[3] = {(cuuint32_t)(128 + 4 * S), 1, 12};
        r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        nfl = (128 + 4 * S) * 12; c[0] = -2 * S; c[1] = 0; c[2] = -2;
    } else if (cs == 4) {  // 3-D float4 plane as (comp, x, y)
        rank = 3; cuuint64_t dims[3] = {4, W, H}, str[2] = {16, (cuuint64_t)Wp * 16}; cuuint32_t box[3] = {4, 132, 12};
        r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        nfl = 4 * 132 * 12; c[0] = 0; c[1] = -2; c[2] = -2;
    } else if (cs == 5 || cs == 6) {  // 4-D float4 plane (comp, x, phase, k)
        S = cs == 5 ? 1 : 4; rank = 4;
        cuuint64_t dims[4] = {4, W, (cuuint64_t)S, (cuuint64_t)(Hp / S)}, str[3] = {16, (cuuint64_t)Wp * 16, (cuuint64_t)Wp * 16 * S};
        cuuint32_t box[4] = {4, (cuuint32_t)(128 + 4 * S), 1, 12};
        r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        nfl = 4 * (128 + 4 * S) * 12; c[0] = 0; c[1] = -2 * S; c[2] = 0; c[3] = -2;
    } else if (cs == 7) {  // float4 plane as 2-D of 16-byte-wide rows? use x dimension in floats, box 256
        rank = 2; cuuint64_t dims[2] = {(cuuint64_t)W * 4, H}, str[1] = {(cuuint64_t)Wp * 16}; cuuint32_t box[2] = {256, 12};
        r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        nfl = 256 * 12; c[0] = -8; c[1] = -2;
    } else if (cs == 8) {  // 3-D with unit-size middle dim replaced: (x, k) 2-D with row stride S*pitch and base offset (per-phase map)
        S = 4; rank = 2; cuuint64_t dims[2] = {W, (cuuint64_t)(Hp / S)}, str[1] = {(cuuint64_t)Wp * 4 * S}; cuuint32_t box[2] = {144, 12};
        r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d + Wp * 1, dims, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        nfl = 144 * 12; c[0] = -8; c[1] = -2;
    } else if (cs == 9) {  // 4-D like case 5 but L2 promotion NONE
        S = 1; rank = 4;
        cuuint64_t dims[4] = {4, W, (cuuint64_t)S, (cuuint64_t)(Hp / S)}, str[3] = {16, (cuuint64_t)Wp * 16, (cuuint64_t)Wp * 16 * S};
        cuuint32_t box[4] = {4, 132, 1, 12};
        r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        nfl = 4 * 132 * 12; c[0] = 0; c[1] = 4; c[2] = 0; c[3] = 4;
    } else { printf("no such case\n"); return 1; }
    if (r != CUDA_SUCCESS) { printf("case %d: encode failed %d\n", cs, (int)r); return 3; }
    int smem = ((nfl * 4 + 127) & ~127) + 16 + 128;
    if (rank == 2) { CK(cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); probe<2><<<1, 128, smem>>>(m, out, nfl, c[0], c[1], c[2], c[3]); }
    if (rank == 3) { CK(cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); probe<3><<<1, 128, smem>>>(m, out, nfl, c[0], c[1], c[2], c[3]); }
    if (rank == 4) { CK(cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); probe<4><<<1, 128, smem>>>(m, out, nfl, c[0], c[1], c[2], c[3]); }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> o(nfl);
    CK(cudaMemcpy(o.data(), out, (size_t)nfl * 4, cudaMemcpyDeviceToHost));
    printf("case %d ok: first values %g %g %g %g ... [%d]=%g\n", cs, o[0], o[1], o[2], o[3], nfl / 2, o[nfl / 2]);
    return 0;
}
