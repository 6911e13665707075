/*
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): a host build of the GPU path front end's
 * per-contour core (rive-runtime_b200/csrc/front_end_core.h, the code the F1 kernels run per
 * thread), driven by three serial loops in place of the three kernels and two prefix scans.
 * tests/test_front_end_cpu.py compares its output byte for byte with the spans / contours / path
 * records the reference front end wrote into the committed flush traces, so the stroke and fill
 * arithmetic is checked on every CPU run of the suite; tests/test_front_end_gpu.py then checks
 * that the kernels produce the same bytes on the device. Nothing in the product links this.
 */
#include "front_end_core.h"

#include <vector>

using namespace rivecuda::fe;

extern "C" int front_end_host_paths(const float* pointsXY,
                                    const uint8_t* verbs,
                                    const rivecuda_path* paths,
                                    uint32_t pathCount,
                                    uint32_t frameWidth, // 0: no frame cull
                                    uint32_t frameHeight,
                                    uint32_t* spans,     // capacity in 64-byte records
                                    uint32_t spanCapacity,
                                    uint32_t* contours,  // 16-byte records
                                    uint32_t* pathData,  // 64-byte records, record 0 reserved
                                    uint32_t* paintData, // 8-byte records
                                    uint32_t* paintAux,  // 128-byte records
                                    rivecuda_front_end_result* result,
                                    // the tables the paths index (rivecuda.h); each may be null
                                    const rivecuda_clip_rect* clipRects,
                                    const rivecuda_gradient_paint* gradientPaints,
                                    const rivecuda_image_paint* imagePaints)
{
    const V2* points = reinterpret_cast<const V2*>(pointsXY);
    std::vector<PathTotals> own(pathCount), prefix(pathCount);
    PathTotals sum = {0, 0, 0, 0};
    for (uint32_t i = 0; i < pathCount; ++i)
    {
        own[i] = count_path(paths[i], points, verbs, frameWidth, frameHeight, clipRects);
        prefix[i] = sum;
        sum.tessVertices += own[i].tessVertices;
        sum.contours += own[i].contours;
        sum.paths += own[i].paths;
    }
    FrontEndOut out;
    out.spans = spans, out.contours = contours, out.pathData = pathData, out.paintData = paintData, out.paintAux = paintAux;
    out.clipRects = clipRects, out.gradientPaints = gradientPaints, out.imagePaints = imagePaints;
    out.spanBase = 0;
    uint32_t padding[2];
    emit_padding_spans(spans, sum.tessVertices, padding);
    out.spanBase = padding[0];
    for (uint32_t i = 0; i < pathCount; ++i)
    {
        prefix[i].spans = sum.spans;
        sum.spans += place_path<false>(paths[i], points, verbs, prefix[i], own[i].tessVertices, out);
    }
    if (out.spanBase + sum.spans > spanCapacity)
        return 1;
    for (uint32_t i = 0; i < pathCount; ++i)
        place_path<true>(paths[i], points, verbs, prefix[i], own[i].tessVertices, out);
    result->path_count = sum.paths + 1;
    result->contour_count = sum.contours;
    result->tess_vertex_span_count = out.spanBase + sum.spans;
    result->midpoint_fan_tess_vertex_count = sum.tessVertices;
    result->tess_data_height = (padding[1] + kTessTextureWidth - 1) / kTessTextureWidth;
    result->first_patch = 1;
    result->patch_count = sum.tessVertices / kPatchSpan;
    result->reserved0 = 0;
    return 0;
}

// Single functions of the core, for known-answer tests that follow the reference's own unit tests
// (tests/unit_tests/runtime/bezier_utils_test.cpp, wangs_formula_test.cpp).
extern "C" int fe_find_cubic_convex_180_chops(const float pts[8], float T[2], int* areCusps)
{
    bool cusps = false;
    const int n = find_cubic_convex_180_chops(reinterpret_cast<const V2*>(pts), T, &cusps);
    *areCusps = cusps ? 1 : 0;
    return n;
}

extern "C" void fe_chop_cubic_at(const float pts[8], float dst[14], float t)
{
    chop_cubic_at(reinterpret_cast<const V2*>(pts), reinterpret_cast<V2*>(dst), t);
}

extern "C" void fe_chop_cubic_at2(const float pts[8], float dst[20], float t0, float t1)
{
    chop_cubic_at(reinterpret_cast<const V2*>(pts), reinterpret_cast<V2*>(dst), t0, t1);
}

extern "C" void fe_eval_cubic_at(const float pts[8], float t, float out[2])
{
    const V2 p = eval_cubic_at(reinterpret_cast<const V2*>(pts), t);
    out[0] = p.x;
    out[1] = p.y;
}

extern "C" uint32_t fe_polar_segments(const float t0[2], const float t1[2], float polarSegmentsPerRadian)
{
    return polar_segments(V2{t0[0], t0[1]}, V2{t1[0], t1[1]}, polarSegmentsPerRadian);
}

extern "C" float fe_fast_acos(float x) { return fast_acos(x); }

extern "C" uint32_t fe_wang_cubic_segments(const float pts[8], const float matrix[6])
{
    return wang_cubic_segments(reinterpret_cast<const V2*>(pts), matrix);
}
