// records.cuh -- the record writer of the fused pipeline: one sorted (tile, Gaussian) intersection -> its entry in
// gaussian_ids_sorted and in the A / B / F record streams the blend kernels stage by TMA (blend.cu).
#pragma once
#include "splat_math.cuh"

struct GfbPackArgs {
    const float2* uv;
    const float* conic;
    const float* opacity;
    const float* feature;
    int C;
    float4* sA;
    float4* sB;
    float4* sF;
    int32_t* ids;
};

__device__ __forceinline__ void gfb_write_record(const GfbPackArgs& a, long long k, int id) {
    const float2 p = a.uv[id];
    const float ca = a.conic[3 * id], cb = a.conic[3 * id + 1], cc = a.conic[3 * id + 2];
    const float o = a.opacity[id];
    float hx, hy;
    gfbm::splat_bbox(ca, cb, cc, o, hx, hy);
    const float* f = a.feature + (size_t)id * a.C;
    float4 fr = make_float4(f[0], 0.0f, 0.0f, 0.0f);
    if (a.C > 1) fr.y = f[1];
    if (a.C > 2) fr.z = f[2];
    if (a.C > 3) fr.w = f[3];
    a.ids[k] = id;
    a.sA[k] = gfb_pack_record_a(p.x, p.y, hx, hy, id);
    a.sB[k] = make_float4(ca, cb, cc, o);
    a.sF[k] = fr;
}
