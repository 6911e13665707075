/*
 * CudaPathRenderer -- a rive::Renderer for frames made only of plain draws (solid colour,
 * unfeathered nonZero / evenOdd / clockwise fills and strokes) that hands the frame's
 * RawPaths to the device instead of running the reference's per-path CPU front end
 * (SURVEY.md 8(f1)). It sits where RiveRenderer sits:
 *
 *     CudaPathRenderer renderer(impl, target, LoadAction::clear, clearColor);
 *     artboardOrScene->draw(&renderer);          // save / restore / transform / drawPath
 *     if (!renderer.flush()) ...                 // rivecuda_front_end_paths + rivecuda_flush
 *
 * What RiveRenderer::drawPath checks before building a PathDraw is mirrored here
 * (rive_renderer.cpp:121-154): empty paths and strokes with !(thickness > 0) are skipped.
 * Per stroked path the two scalars PathDraw computes with libm are computed here the same way
 * (draw.cpp:603-607, 776-813), and a modulated opacity goes into the colour as PathDraw puts it
 * there (draw.cpp:727-737); blend modes travel in the paint record; clip RECTANGLES (clipPath with an
 * axis-aligned rectangle, nested ones intersected) travel as a table the paths index. Anything else -- clip paths, gradients, images, feathers --
 * is not handled by the device front end: the renderer records the first such call
 * and flush() refuses the frame, so the caller can draw it with RiveRenderer (no silent fallback).
 */
#pragma once

#include "render_context_cuda_impl.hpp"

#include "rive/math/bezier_utils.hpp"
#include "rive/math/mat2d.hpp"
#include "rive/renderer.hpp"
#include "rive/renderer/gpu.hpp"
#include "rive/shapes/paint/color.hpp"
#include "rive_render_paint.hpp"
#include "rive_render_path.hpp"
#include "rive/renderer/rive_renderer.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <string>
#include <vector>

namespace rive::gpu
{
class CudaPathRenderer : public Renderer
{
public:
    CudaPathRenderer(RenderContextCUDAImpl* impl, RenderTargetCUDA* target, LoadAction loadAction, ColorInt clearColor) :
        m_impl(impl), m_target(target), m_loadAction(loadAction), m_clearColor(clearColor)
    {}

    void save() override { m_stack.push_back(m_stack.back()); }
    void restore() override
    {
        if (m_stack.size() > 1)
            m_stack.pop_back();
    }
    void transform(const Mat2D& m) override { m_stack.back().matrix = m_stack.back().matrix * m; }
    // RiveRenderer::modulateOpacity (rive_renderer.cpp:115-119): part of the save / restore state.
    void modulateOpacity(float opacity) override { m_stack.back().opacity = std::max(0.0f, m_stack.back().opacity * opacity); }

    void drawPath(RenderPath* renderPath, RenderPaint* renderPaint) override
    {
        auto* path = static_cast<RiveRenderPath*>(renderPath);
        auto* paint = static_cast<RiveRenderPaint*>(renderPaint);
        const RawPath& raw = path->getRawPath();
        if (raw.empty() || (paint->getIsStroked() && !(paint->getThickness() > 0)) || !(paint->getFeather() >= 0))
            return;
        if (paint->getFeather() != 0)
            return refuse("drawPath with a feather");
        if (paint->getType() != PaintType::solidColor)
            return refuse("drawPath with a gradient");
        if (paint->getImageTexture() != nullptr)
            return refuse("drawPath with an image paint");
        if (m_stack.back().overallClipPixelBounds.empty())
            return; // rive_renderer.cpp:151
        const Mat2D& m = m_stack.back().matrix;
        rivecuda_path p;
        memset(&p, 0, sizeof(p));
        p.first_verb = static_cast<uint32_t>(m_verbs.size());
        p.verb_count = static_cast<uint32_t>(raw.verbs().size());
        p.first_point = static_cast<uint32_t>(m_points.size());
        for (int i = 0; i < 6; ++i)
            p.matrix[i] = m[i];
        // PathDraw applies the modulated opacity to a solid colour (draw.cpp:727-737).
        p.blend_mode = ConvertBlendModeToPLSBlendMode(paint->getBlendMode()); // PaintData::set (gpu.cpp:889)
        p.color = m_stack.back().opacity != 1.0f ? colorModulateOpacity(paint->getColor(), m_stack.back().opacity) : paint->getColor();
        p.stroke = m_stack.back().clipRectIndex << 8; // 0: no clip rectangle
        if (paint->getIsStroked())
        {
            p.stroke |= 1;
            p.stroke_radius = fmaxf(paint->getThickness() * .5f, FLT_MIN); // draw.cpp:603-607
            p.join = static_cast<uint32_t>(paint->getJoin());
            p.cap = static_cast<uint32_t>(paint->getCap());
            p.matrix_max_scale = m.findMaxScale(); // draw.cpp:778
            p.polar_segments_per_radian =
                math::calc_polar_segments_per_radian<kPolarPrecision>(p.stroke_radius * p.matrix_max_scale); // draw.cpp:806-808
        }
        else
        {
            // clockwise: the device front end picks the contour directions from the matrix
            // (draw.cpp:657-680) and the host gives its batch ShaderMiscFlags::clockwiseFill.
            p.fill_rule = path->getFillRule() == FillRule::evenOdd ? 1 : path->getFillRule() == FillRule::clockwise ? 2 : 0;
        }
        for (PathVerb v : raw.verbs())
            m_verbs.push_back(static_cast<uint8_t>(v));
        m_points.insert(m_points.end(), raw.points().begin(), raw.points().end());
        m_paths.push_back(p);
    }

    // Clip rectangles (the ENABLE_CLIP_RECT feature): what RiveRenderer::clipPath / clipRectImpl do
    // with an axis-aligned rectangle (rive_renderer.cpp:199-322). Any other clip is a clip PATH
    // (stencil-like updates of the clip plane between the draws), which this renderer refuses.
    void clipPath(RenderPath* renderPath) override
    {
        auto* path = static_cast<RiveRenderPath*>(renderPath);
        State& state = m_stack.back();
        if (state.overallClipPixelBounds.empty())
            return;
        if (path->getRawPath().empty())
        {
            state.overallClipPixelBounds = {};
            return;
        }
        AABB rect;
        if (!RiveRenderer::IsAABB(path->getRawPath(), &rect))
        {
            refuse("clipPath with something other than an axis-aligned rectangle");
            return;
        }
        if (rect.isEmptyOrNaN())
        {
            state.overallClipPixelBounds = {};
            return;
        }
        if (state.hasClipRect && !(state.matrix == state.clipRectMatrix))
        {
            // A second rectangle only intersects with the first in the first one's space, and
            // only if it is still a rectangle there (transform_rect_to_new_space).
            Mat2D currentToNew;
            if (!state.clipRectMatrix.invert(&currentToNew))
            {
                refuse("clipPath: nested clip rectangle under a singular matrix");
                return;
            }
            currentToNew = currentToNew * state.matrix;
            const float maxSkew = fmaxf(fabsf(currentToNew.xy()), fabsf(currentToNew.yx()));
            const float maxScale = fmaxf(fabsf(currentToNew.xx()), fabsf(currentToNew.yy()));
            if (maxSkew > math::EPSILON && maxScale > math::EPSILON)
            {
                refuse("clipPath: nested clip rectangle that is not axis-aligned with the first");
                return;
            }
            Vec2D pts[2] = {{rect.left(), rect.top()}, {rect.right(), rect.bottom()}};
            currentToNew.mapPoints(pts, pts, 2);
            rect = {std::min(pts[0].x, pts[1].x), std::min(pts[0].y, pts[1].y), std::max(pts[0].x, pts[1].x), std::max(pts[0].y, pts[1].y)};
        }
        if (!state.hasClipRect)
        {
            state.clipRect = rect;
            state.clipRectMatrix = state.matrix;
            state.hasClipRect = true;
        }
        else
        {
            state.clipRect = {std::max(state.clipRect.left(), rect.left()), std::max(state.clipRect.top(), rect.top()),
                              std::min(state.clipRect.right(), rect.right()), std::min(state.clipRect.bottom(), rect.bottom())};
        }
        const IAABB clipRectPixelBounds = state.clipRectMatrix.mapBoundingBox(state.clipRect).roundOut();
        state.overallClipPixelBounds = state.overallClipPixelBounds.intersect(clipRectPixelBounds);
        // The record every draw under this state carries (Draw::setClipRect + PaintAuxData::set).
        const ClipRectInverseMatrix inverse(state.clipRectMatrix, state.clipRect);
        const Mat2D& m = inverse.inverseMatrix();
        rivecuda_clip_rect record;
        for (int i = 0; i < 6; ++i)
            record.inverse_matrix[i] = m[i];
        record.inverse_fwidth[0] = -1.f / (fabsf(m.xx()) + fabsf(m.xy())); // gpu.cpp:1052-1053
        record.inverse_fwidth[1] = -1.f / (fabsf(m.yx()) + fabsf(m.yy()));
        // (The bounds only change in this function, so they are the state's at every draw under it.)
        record.pixel_bounds[0] = state.overallClipPixelBounds.left;
        record.pixel_bounds[1] = state.overallClipPixelBounds.top;
        record.pixel_bounds[2] = state.overallClipPixelBounds.right;
        record.pixel_bounds[3] = state.overallClipPixelBounds.bottom;
        m_clipRects.push_back(record);
        state.clipRectIndex = static_cast<uint32_t>(m_clipRects.size());
    }
    void drawImage(const RenderImage*, ImageSampler, BlendMode, float) override { refuse("drawImage"); }
    void drawImageMesh(const RenderImage*,
                       ImageSampler,
                       rcp<RenderBuffer>,
                       rcp<RenderBuffer>,
                       rcp<RenderBuffer>,
                       uint32_t,
                       uint32_t,
                       BlendMode,
                       float) override
    {
        refuse("drawImageMesh");
    }

    // The first call this renderer cannot express, or nullptr.
    const char* refusedCall() const { return m_refused.empty() ? nullptr : m_refused.c_str(); }
    size_t pathCount() const { return m_paths.size(); }

    // Renders the collected frame. False if a call was refused or the device reports an error.
    bool flush()
    {
        if (!m_refused.empty())
        {
            fprintf(stderr, "CudaPathRenderer: the frame contains %s; draw it with RiveRenderer\n", m_refused.c_str());
            return false;
        }
        RenderContextCUDAImpl::PlainPathFrame frame;
        frame.renderTarget = m_target;
        frame.loadAction = m_loadAction;
        frame.clearColor = m_clearColor;
        frame.points = m_points.data();
        frame.pointCount = m_points.size();
        frame.verbs = m_verbs.data();
        frame.verbCount = m_verbs.size();
        frame.paths = m_paths.data();
        frame.pathCount = m_paths.size();
        frame.clipRects = m_clipRects.data();
        frame.clipRectCount = m_clipRects.size();
        return m_impl->flushPlainPaths(frame);
    }

private:
    void refuse(const char* what)
    {
        if (m_refused.empty())
            m_refused = what;
        if (getenv("RIVECUDA_FRONT_END_VERBOSE") != nullptr)
            fprintf(stderr, "CudaPathRenderer: refused %s\n", what);
    }

    RenderContextCUDAImpl* m_impl;
    RenderTargetCUDA* m_target;
    LoadAction m_loadAction;
    ColorInt m_clearColor;
    struct State
    {
        Mat2D matrix;
        float opacity = 1.0f;
        // RiveRenderer::RenderState's clip-rectangle members (rive_renderer.hpp:100-110).
        IAABB overallClipPixelBounds = IAABB::makeMaximal();
        bool hasClipRect = false;
        AABB clipRect;
        Mat2D clipRectMatrix;
        uint32_t clipRectIndex = 0; // 1 + index into m_clipRects of the state's rectangle
    };
    std::vector<State> m_stack{State()};
    std::vector<Vec2D> m_points;
    std::vector<uint8_t> m_verbs;
    std::vector<rivecuda_path> m_paths;
    std::vector<rivecuda_clip_rect> m_clipRects;
    std::string m_refused;
};
} // namespace rive::gpu
