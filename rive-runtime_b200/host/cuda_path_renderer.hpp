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
 * axis-aligned rectangle, nested ones intersected) travel as a table the paths index; linear and
 * radial gradients get their colour ramps allocated here as LogicalFlush::allocateGradient does, and
 * travel as GradientSpans plus a table of paint records; clip PATHS become clipUpdate paths in
 * front of the draws that need them, under the clip IDs RiveRenderer::applyClip would hand out;
 * image paints and drawImage travel as a table of image matrices, their textures on the batches.
 * Image meshes are passed through as batches of their own between the paths'. Anything else -- feathers --
 * is not handled by the device front end: the renderer records the first such call
 * and flush() refuses the frame, so the caller can draw it with RiveRenderer (no silent fallback).
 */
#pragma once

#include "render_context_cuda_impl.hpp"
#include "../csrc/front_end_core.h" // the frame cull the device applies (fe::is_outside_frame), host build

#include "rive/math/bezier_utils.hpp"
#include "rive/math/mat2d.hpp"
#include "rive/renderer.hpp"
#include "rive/renderer/gpu.hpp"
#include "rive/shapes/paint/color.hpp"
#include "gradient.hpp"
#include "rive_render_paint.hpp"
#include "rive_render_path.hpp"
#include "rive/renderer/rive_renderer.hpp"
#include "rive/renderer/rive_render_image.hpp"

#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace rive::gpu
{
class CudaPathRenderer : public Renderer
{
public:
    CudaPathRenderer(RenderContextCUDAImpl* impl, RenderTargetCUDA* target, LoadAction loadAction, ColorInt clearColor) :
        m_impl(impl), m_target(target), m_loadAction(loadAction), m_clearColor(clearColor)
    {}

    // Draws this renderer cannot express (feathers) can be DELEGATED to the reference's own front
    // end instead of refusing the frame: what has been collected so far is flushed, the draw goes
    // through a RiveRenderer on the owning RenderContext into the same target (a flush of its own,
    // LoadAction::preserveRenderTarget), and collecting goes on on top of it -- a frame may be cut
    // into flushes anywhere, each composites in order (the reference cuts frames into logical
    // flushes itself). begin(loadAction, clearColor) begins a frame on the RenderContext and returns
    // a RiveRenderer for it; end() flushes it into the target. Without hooks such a frame is refused.
    using BeginReferenceFrame = std::function<std::unique_ptr<Renderer>(LoadAction, ColorInt)>;
    using EndReferenceFrame = std::function<void()>;
    void setDelegate(BeginReferenceFrame begin, EndReferenceFrame end)
    {
        m_beginReference = std::move(begin);
        m_endReference = std::move(end);
    }
    size_t delegatedDrawCount() const { return m_delegatedDraws; }
    // Optional: also delegate the fills the reference would draw by INTERIOR TRIANGULATION
    // (TriangulationController::isEligible, triangulation_controller.hpp:70-77: >= 512 x 512 px under
    // the default thresholds, at most 256 verbs). The device front end draws every fill as a midpoint
    // fan, which is inside the reference's tessellation tolerance but not the same pixels; with this
    // on, such fills go through the reference's triangulator (more flushes per frame, the reference's
    // default pixels).
    void setLargeFillDelegation(const TriangulationThresholds& thresholds)
    {
        m_delegateLargeFills = true;
        m_thresholds = thresholds;
    }

    // save / restore / transform / clipPath / modulateOpacity are also kept as the calls they were,
    // per open save() scope: replayed into a fresh RiveRenderer they rebuild its state with the very
    // float operations that built it here (a closed scope leaves nothing behind).
    void save() override
    {
        m_stack.push_back(m_stack.back());
        m_scopes.emplace_back();
        if (m_reference != nullptr)
            m_reference->save();
    }
    void restore() override
    {
        if (m_stack.size() > 1)
        {
            m_stack.pop_back();
            m_scopes.pop_back();
            if (m_reference != nullptr)
                m_reference->restore();
        }
    }
    void transform(const Mat2D& m) override
    {
        m_stack.back().matrix = m_stack.back().matrix * m;
        StateCall call;
        call.kind = StateCall::Transform;
        call.matrix = m;
        m_scopes.back().push_back(call);
        if (m_reference != nullptr)
            m_reference->transform(m);
    }
    // RiveRenderer::modulateOpacity (rive_renderer.cpp:115-119): part of the save / restore state.
    void modulateOpacity(float opacity) override
    {
        m_stack.back().opacity = std::max(0.0f, m_stack.back().opacity * opacity);
        StateCall call;
        call.kind = StateCall::ModulateOpacity;
        call.opacity = opacity;
        m_scopes.back().push_back(call);
        if (m_reference != nullptr)
            m_reference->modulateOpacity(opacity);
    }

    void drawPath(RenderPath* renderPath, RenderPaint* renderPaint) override
    {
        // (null or foreign objects are ignored, as RiveRenderer's LITE_RTTI_CAST_OR_RETURN does)
        auto* path = lite_rtti_cast<RiveRenderPath*>(renderPath);
        auto* paint = lite_rtti_cast<RiveRenderPaint*>(renderPaint);
        if (path == nullptr || paint == nullptr)
            return;
        const RawPath& raw = path->getRawPath();
        if (raw.empty() || (paint->getIsStroked() && !(paint->getThickness() > 0)) || !(paint->getFeather() >= 0))
            return;
        if (paint->getFeather() != 0)
        {
            // RiveRenderer::drawPath does not draw feathered fills that are not clockwise
            // (rive_renderer.cpp:163-171): no need to open a reference frame for those.
            if (!paint->getIsStroked() && path->getFillRule() != FillRule::clockwise)
                return;
            if (m_stack.back().overallClipPixelBounds.empty())
                return;
            if (!m_beginReference)
                return refuse("drawPath with a feather");
            openReference();
            m_reference->drawPath(renderPath, renderPaint);
            ++m_delegatedDraws;
            return;
        }
        if (m_delegateLargeFills && m_beginReference && !paint->getIsStroked() && !m_stack.back().overallClipPixelBounds.empty())
        {
            const float area = find_transformed_area(path->getBounds(), m_stack.back().matrix); // draw.cpp:518-520
            if (m_thresholds.frameBudgetMs > 0 && area >= m_thresholds.minArea && raw.verbs().count() <= m_thresholds.maxVerbs)
            {
                openReference();
                m_reference->drawPath(renderPath, renderPaint);
                ++m_delegatedDraws;
                return;
            }
        }
        closeReference(); // (a draw of our own: what was delegated before it is flushed first)
        if (paint->getType() != PaintType::solidColor && paint->getType() != PaintType::linearGradient && paint->getType() != PaintType::radialGradient)
            return refuse("drawPath with an unknown paint type");
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
        const State& state = m_stack.back();
        if (state.clipStackHeight != 0 || paint->getType() != PaintType::solidColor)
        {
            // A draw that PathDraw::Make (draw.cpp:439-509) or RiveRenderer::applyClip
            // (rive_renderer.cpp:636-646) culls allocates no colour ramp and triggers no clip update:
            // apply the cull the device would apply (the same code, built for the host) first.
            static_assert(sizeof(Vec2D) == sizeof(rivecuda::fe::V2), "points are passed as they are");
            rivecuda_clip_rect bounds;
            memset(&bounds, 0, sizeof(bounds));
            bounds.pixel_bounds[0] = state.overallClipPixelBounds.left;
            bounds.pixel_bounds[1] = state.overallClipPixelBounds.top;
            bounds.pixel_bounds[2] = state.overallClipPixelBounds.right;
            bounds.pixel_bounds[3] = state.overallClipPixelBounds.bottom;
            rivecuda_path probe = p;
            probe.stroke = (p.stroke & 1u) | (1u << 8);
            if (rivecuda::fe::is_outside_frame(probe, reinterpret_cast<const rivecuda::fe::V2*>(raw.points().data()), static_cast<uint32_t>(raw.points().size()),
                                               m_target->width(), m_target->height(), &bounds))
                return;
        }
        GradientDraw gradientDraw;
        if (paint->getType() != PaintType::solidColor)
        {
            // PathDraw keeps the gradient with the modulated opacity folded into its colours
            // (draw.cpp:580) and allocates its colour ramp when the draw is pushed. When the gradient
            // texture is full the reference starts a new logical flush and pushes the draw again
            // (RiveRenderer::clipAndPushDraw, rive_renderer.cpp:508-556): so does this renderer.
            // (The ramp is allocated before the draw's clip updates are emitted -- they have no
            // ramps, so the order among ramps is the reference's -- which leaves nothing to undo.)
            gradientDraw.gradient = paint->getGradientWithOpacity(m_stack.back().opacity);
            gradientDraw.matrix = m;
            if (gradientDraw.gradient == nullptr)
                return refuse("drawPath with a gradient paint that has no gradient");
            if (!allocateGradient(gradientDraw.gradient.get(), &gradientDraw.location))
            {
                flushAndContinue();
                if (!allocateGradient(gradientDraw.gradient.get(), &gradientDraw.location))
                    return refuse("drawPath with a gradient that does not fit an empty gradient texture");
            }
        }
        if (state.clipStackHeight != 0)
        {
            const uint32_t clipID = applyClip(state.clipStackHeight);
            if (clipID == 0)
                return refuse("drawPath under more clip updates than one flush has clip IDs");
            p.blend_mode |= clipID << 16;
        }
        // (applyClip and a flush in between move where this path's verbs and points go)
        p.first_verb = static_cast<uint32_t>(m_verbs.size());
        p.first_point = static_cast<uint32_t>(m_points.size());
        if (paint->getImageTexture() != nullptr)
        {
            // An image paint (RenderPaint::modulatedImage, or drawImage below): the words
            // PaintAuxData::set computes for it (gpu.cpp:1001-1033), by the reference's own writer.
            const Mat2D imageMatrix = m * paint->getImageTransform(); // rive_renderer.cpp:156-162
            PaintAuxData aux;
            aux.set(m, imageMatrix, PaintType::solidColor, SimplePaintValue(), nullptr, paint->getImageTexture(), nullptr, m_target, m_impl->platformFeatures());
            float words[32];
            memcpy(words, &aux, sizeof(words));
            rivecuda_image_paint record;
            memcpy(record.image_matrix, words + 16, 24);
            record.image_texture_lod = words[22];
            record.reserved0 = 0;
            m_imagePaints.push_back(record);
            RenderContextCUDAImpl::PlainImageBinding binding;
            binding.texture = static_cast<const TextureCUDA*>(paint->getImageTexture())->handle();
            binding.samplerKey = paint->getImageSampler().asKey();
            m_imageBindings.push_back(binding);
            m_imageTextures.push_back(ref_rcp(paint->getImageTexture()));
            p.cap |= static_cast<uint32_t>(m_imagePaints.size()) << 8;
        }
        if (paint->getType() != PaintType::solidColor)
        {
            m_gradientDraws.push_back(std::move(gradientDraw));
            p.fill_rule |= static_cast<uint32_t>(m_gradientDraws.size()) << 8;
        }
        for (PathVerb v : raw.verbs())
            m_verbs.push_back(static_cast<uint8_t>(v));
        m_points.insert(m_points.end(), raw.points().begin(), raw.points().end());
        m_paths.push_back(p);
    }

    // Clip rectangles (the ENABLE_CLIP_RECT feature): what RiveRenderer::clipPath / clipRectImpl do
    // with an axis-aligned rectangle (rive_renderer.cpp:199-322). Any other clip is a clip PATH
    // (updates of the clip plane between the draws): clipPathImpl / applyClip below.
    void clipPath(RenderPath* renderPath) override
    {
        auto* path = lite_rtti_cast<RiveRenderPath*>(renderPath);
        if (path == nullptr)
            return;
        {
            StateCall call;
            call.kind = StateCall::ClipPath;
            call.path = ref_rcp(renderPath);
            m_scopes.back().push_back(std::move(call));
            if (m_reference != nullptr)
                m_reference->clipPath(renderPath);
        }
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
            return clipPathImpl(path);
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
                return clipPathImpl(path); // not a rectangle in the first one's space: a clip path
            currentToNew = currentToNew * state.matrix;
            const float maxSkew = fmaxf(fabsf(currentToNew.xy()), fabsf(currentToNew.yx()));
            const float maxScale = fmaxf(fabsf(currentToNew.xx()), fabsf(currentToNew.yy()));
            if (maxSkew > math::EPSILON && maxScale > math::EPSILON)
                return clipPathImpl(path);
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
    // RiveRenderer::drawImage (rive_renderer.cpp:383-446): the unit rectangle under a matrix scaled
    // by the image size, with an image paint. (The opacity goes through drawPath's modulation again,
    // as it does in the reference.)
    void drawImage(const RenderImage* renderImage, ImageSampler sampler, BlendMode blendMode, float opacity) override
    {
        auto* image = lite_rtti_cast<const RiveRenderImage*>(renderImage);
        if (image == nullptr)
            return;
        rcp<Texture> texture = image->refTexture();
        if (texture == nullptr)
            return;
        const float finalOpacity = std::max(0.0f, opacity * m_stack.back().opacity);
        save();
        transform(Mat2D::fromScale(static_cast<float>(image->width()), static_cast<float>(image->height())));
        if (m_unitRectPath == nullptr)
        {
            m_unitRectPath = make_rcp<RiveRenderPath>();
            m_unitRectPath->line({1, 0});
            m_unitRectPath->line({1, 1});
            m_unitRectPath->line({0, 1});
        }
        RiveRenderPaint paint;
        paint.image(std::move(texture), finalOpacity);
        paint.blendMode(blendMode);
        paint.imageSampler(sampler);
        drawPath(m_unitRectPath.get(), &paint);
        restore();
    }
    // RiveRenderer::drawImageMesh (rive_renderer.cpp:448-494): an ImageMeshDraw over the whole render
    // target, clipped like any other draw (applyClip); the backend draws it between the path batches.
    void drawImageMesh(const RenderImage* renderImage,
                       ImageSampler sampler,
                       rcp<RenderBuffer> vertices,
                       rcp<RenderBuffer> uvCoords,
                       rcp<RenderBuffer> indices,
                       uint32_t,
                       uint32_t indexCount,
                       BlendMode blendMode,
                       float opacity) override
    {
        auto* image = lite_rtti_cast<const RiveRenderImage*>(renderImage);
        if (image == nullptr)
            return;
        rcp<Texture> texture = image->refTexture();
        const State& state = m_stack.back();
        if (texture == nullptr || state.overallClipPixelBounds.empty())
            return;
        closeReference();
        RenderContextCUDAImpl::PlainMeshDraw mesh;
        mesh.vertexBuffer = RenderContextCUDAImpl::renderBufferHandle(vertices.get());
        mesh.uvBuffer = RenderContextCUDAImpl::renderBufferHandle(uvCoords.get());
        mesh.indexBuffer = RenderContextCUDAImpl::renderBufferHandle(indices.get());
        if (mesh.vertexBuffer == nullptr || mesh.uvBuffer == nullptr || mesh.indexBuffer == nullptr)
            return; // foreign buffers: skipped, like LITE_RTTI_CAST_OR_BREAK in the backends
        const float finalOpacity = std::max(0.0f, opacity * state.opacity);
        // applyClip: the draw's bounds are the whole target, clipped by the state's.
        const IAABB clipped = state.overallClipPixelBounds.intersect(Draw::FULLSCREEN_PIXEL_BOUNDS);
        if (clipped.empty() || clipped.left >= static_cast<int32_t>(m_target->width()) || clipped.top >= static_cast<int32_t>(m_target->height()) ||
            clipped.right <= 0 || clipped.bottom <= 0)
            return;
        if (state.clipStackHeight != 0)
        {
            mesh.clipID = applyClip(state.clipStackHeight);
            if (mesh.clipID == 0)
                return refuse("drawImageMesh under more clip updates than one flush has clip IDs");
        }
        mesh.afterPath = m_paths.size();
        mesh.hasClipRect = state.hasClipRect;
        const ClipRectInverseMatrix clipRectInverse = state.hasClipRect ? ClipRectInverseMatrix(state.clipRectMatrix, state.clipRect) : ClipRectInverseMatrix::WideOpen();
        mesh.instance = ImageDrawInstance(state.matrix, finalOpacity, state.hasClipRect ? &clipRectInverse : nullptr, mesh.clipID, blendMode, 0);
        mesh.texture = static_cast<const TextureCUDA*>(texture.get())->handle();
        mesh.samplerKey = sampler.asKey();
        mesh.indexCount = indexCount;
        mesh.blendMode = ConvertBlendModeToPLSBlendMode(blendMode);
        m_meshDraws.push_back(mesh);
        m_imageTextures.push_back(std::move(texture));
        m_meshBuffers.push_back(std::move(vertices));
        m_meshBuffers.push_back(std::move(uvCoords));
        m_meshBuffers.push_back(std::move(indices));
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
        if (m_reference != nullptr)
        {
            // the frame ends with delegated draws: they are flushed, nothing of ours is left to draw
            closeReference();
            return !m_flushFailed;
        }
        return flushCollected() && !m_flushFailed;
    }

private:
    // Delegation (setDelegate): opens a reference frame behind what has been collected so far and
    // brings its RiveRenderer to this renderer's state; consecutive delegated draws share the frame.
    void openReference()
    {
        if (m_reference != nullptr)
            return;
        if (!m_paths.empty() || !m_meshDraws.empty())
            flushAndContinue();
        m_reference = m_beginReference(m_loadAction, m_clearColor);
        for (size_t scope = 0; scope < m_scopes.size(); ++scope)
        {
            if (scope != 0)
                m_reference->save();
            for (const StateCall& call : m_scopes[scope])
            {
                switch (call.kind)
                {
                    case StateCall::Transform:
                        m_reference->transform(call.matrix);
                        break;
                    case StateCall::ClipPath:
                        m_reference->clipPath(call.path.get());
                        break;
                    case StateCall::ModulateOpacity:
                        m_reference->modulateOpacity(call.opacity);
                        break;
                }
            }
        }
    }
    void closeReference()
    {
        if (m_reference == nullptr)
            return;
        m_reference.reset();
        m_endReference();
        // The reference's flush rendered its own clips and ramps: nothing of ours survives it.
        m_clipContentID = 0;
        m_clipCount = 0;
        for (ClipElement& clip : m_clipStack)
            clip.clipID = 0;
        m_loadAction = LoadAction::preserveRenderTarget;
    }

    // What RenderContext::logicalFlush() is to RiveRenderer: draws what has been collected so far
    // and goes on collecting on top of it (the frame ran out of something one flush holds: gradient
    // texture rows; or a draw is being delegated). Colour ramps, image bindings and clip IDs do not
    // outlive a flush: the clip stack's elements are rendered into the clip plane again when the
    // next draw needs them.
    void flushAndContinue()
    {
        if (!flushCollected())
            m_flushFailed = true;
        m_points.clear();
        m_verbs.clear();
        m_paths.clear();
        m_gradientDraws.clear();
        m_simpleGradients.clear();
        m_simpleRamps.clear();
        m_complexGradients.clear();
        m_complexRamps.clear();
        m_imagePaints.clear();
        m_imageBindings.clear();
        m_imageTextures.clear();
        m_meshDraws.clear();
        m_meshBuffers.clear();
        m_clipContentID = 0;
        m_clipCount = 0;
        for (ClipElement& clip : m_clipStack)
            clip.clipID = 0;
        m_loadAction = LoadAction::preserveRenderTarget;
    }

    bool flushCollected()
    {
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
        frame.hasClipPaths = m_hasClipPaths;
        frame.imagePaints = m_imagePaints.data();
        frame.imageBindings = m_imageBindings.data();
        frame.imagePaintCount = m_imagePaints.size();
        frame.meshDraws = m_meshDraws.data();
        frame.meshDrawCount = m_meshDraws.size();
        std::vector<GradientSpan> gradSpans;
        std::vector<rivecuda_gradient_paint> gradientPaints;
        if (!m_gradientDraws.empty())
        {
            // LogicalFlush::layoutResources (render_context.cpp:1206, 1338-1339, 1442-1443): the simple
            // ramps fill the first rows, 256 to a row, one row per complex ramp follows; the paints
            // address rows normalised by the ALLOCATED texture height.
            const uint32_t complexOffsetY = static_cast<uint32_t>((m_simpleRamps.size() + kGradTextureWidthInSimpleRamps - 1) / kGradTextureWidthInSimpleRamps);
            frame.gradDataHeight = complexOffsetY + static_cast<uint32_t>(m_complexRamps.size());
            const uint32_t allocatedHeight = m_impl->reservePlainGradientRows(frame.gradDataHeight);
            const GradTextureLayout layout = {complexOffsetY, 1.f / static_cast<float>(allocatedHeight)};
            writeGradientSpans(complexOffsetY, &gradSpans);
            gradientPaints.reserve(m_gradientDraws.size());
            for (const GradientDraw& draw : m_gradientDraws)
            {
                // The reference's own record writers, on the stack; the words the device copies.
                SimplePaintValue value;
                value.colorRampLocation = draw.location;
                PaintData paintData;
                paintData.set(DrawContents::none, draw.gradient->paintType(), value, layout, 0, false, false, BlendMode::srcOver);
                PaintAuxData aux;
                aux.set(draw.matrix, Mat2D(), draw.gradient->paintType(), value, draw.gradient.get(), nullptr, nullptr, m_target, m_impl->platformFeatures());
                uint32_t paintWords[2];
                float auxWords[8];
                memcpy(paintWords, &paintData, sizeof(paintWords));
                memcpy(auxWords, &aux, sizeof(auxWords));
                rivecuda_gradient_paint record;
                record.paint_type = paintWords[0] & 0xfu;
                memcpy(&record.grad_texture_y, &paintWords[1], 4);
                memcpy(record.paint_matrix, auxWords, 24);
                record.grad_horizontal_span[0] = auxWords[6];
                record.grad_horizontal_span[1] = auxWords[7];
                gradientPaints.push_back(record);
            }
            frame.gradSpans = gradSpans.data();
            frame.gradSpanCount = gradSpans.size();
            frame.gradientPaints = gradientPaints.data();
            frame.gradientPaintCount = gradientPaints.size();
        }
        return m_impl->flushPlainPaths(frame);
    }

private:
    // RiveRenderer::clipPathImpl (rive_renderer.cpp:322-381): the clip stack is shared by all states
    // (a state holds its height), so that a path clipped again after a restore() reuses its element
    // -- and the clip it may still have in the clip plane.
    void clipPathImpl(const RiveRenderPath* path)
    {
        State& state = m_stack.back();
        if (path->getBounds().isEmptyOrNaN())
        {
            state.overallClipPixelBounds = {};
            return;
        }
        const size_t height = state.clipStackHeight;
        if (m_clipStack.size() == height || !m_clipStack[height].isEquivalent(state.matrix, path))
        {
            const IAABB pixelBounds = state.matrix.mapBoundingBox(path->getRawPath().points()).roundOut();
            state.overallClipPixelBounds = state.overallClipPixelBounds.intersect(pixelBounds);
            if (state.overallClipPixelBounds.empty())
                return;
            m_clipStack.resize(height);
            ClipElement element;
            element.matrix = state.matrix;
            element.path = ref_rcp(path);
            element.rawPathMutationID = path->getRawPathMutationID();
            element.fillRule = path->getFillRule();
            element.pixelBounds = pixelBounds;
            m_clipStack.push_back(std::move(element));
        }
        else
        {
            state.overallClipPixelBounds = state.overallClipPixelBounds.intersect(m_clipStack[height].pixelBounds);
            if (state.overallClipPixelBounds.empty())
                return;
        }
        state.clipStackHeight = height + 1;
        m_hasClipPaths = true;
    }

    // RiveRenderer::applyClip in rasterOrdering mode (rive_renderer.cpp:648-822): every clip element
    // above the one currently in the clip plane is drawn as a PaintType::clipUpdate path, nested in
    // its predecessor, under a fresh clip ID. Returns the clip ID the draw is clipped against.
    uint32_t applyClip(size_t clipStackHeight)
    {
        size_t current = static_cast<size_t>(-1);
        if (m_clipContentID != 0)
        {
            for (size_t i = clipStackHeight - 1; i != static_cast<size_t>(-1); --i)
            {
                if (m_clipStack[i].clipID == m_clipContentID)
                {
                    current = i;
                    break;
                }
            }
        }
        uint32_t parentClipID = current == static_cast<size_t>(-1) ? 0u : m_clipStack[current].clipID;
        for (size_t i = current + 1; i < clipStackHeight; ++i)
        {
            ClipElement& clip = m_clipStack[i];
            if (m_clipCount >= kMaxClipID)
                return 0;
            clip.clipID = ++m_clipCount; // LogicalFlush::generateClipID (render_context.cpp:480-492)
            const RawPath& raw = clip.path->getRawPath();
            rivecuda_path c;
            memset(&c, 0, sizeof(c));
            c.first_verb = static_cast<uint32_t>(m_verbs.size());
            c.verb_count = static_cast<uint32_t>(raw.verbs().size());
            c.first_point = static_cast<uint32_t>(m_points.size());
            for (int k = 0; k < 6; ++k)
                c.matrix[k] = clip.matrix[k];
            c.fill_rule = clip.fillRule == FillRule::evenOdd ? 1 : clip.fillRule == FillRule::clockwise ? 2 : 0;
            c.color = parentClipID;                      // RiveRenderPaint::clipUpdate(outerClipID)
            c.blend_mode = 0x100u | (clip.clipID << 16); // Draw::setClipID
            for (PathVerb v : raw.verbs())
                m_verbs.push_back(static_cast<uint8_t>(v));
            m_points.insert(m_points.end(), raw.points().begin(), raw.points().end());
            m_paths.push_back(c);
            parentClipID = clip.clipID;
        }
        m_clipContentID = parentClipID;
        return parentClipID;
    }

    // LogicalFlush::allocateGradient (render_context.cpp:588-674): two-stop 0..1 (and one-stop)
    // gradients share two-texel ramps keyed by their colours, everything else gets a row of its own,
    // shared by gradients of equal content.
    bool allocateGradient(const Gradient* gradient, ColorRampLocation* location)
    {
        const float* stops = gradient->stops();
        const ColorInt* colors = gradient->colors();
        const size_t stopCount = gradient->count();
        auto data_height = [](size_t simple, size_t complex) { return (simple + kGradTextureWidthInSimpleRamps - 1) / kGradTextureWidthInSimpleRamps + complex; };
        if (stopCount == 1 || (stopCount == 2 && stops[0] == 0 && stops[1] == 1))
        {
            const ColorInt ramp[2] = {colors[0], colors[std::min<size_t>(1, stopCount - 1)]};
            const uint64_t key = (static_cast<uint64_t>(ramp[1]) << 32) | ramp[0];
            uint32_t texelIndex;
            auto it = m_simpleGradients.find(key);
            if (it != m_simpleGradients.end())
            {
                texelIndex = it->second;
            }
            else
            {
                if (data_height(m_simpleRamps.size() + 1, m_complexRamps.size()) > RenderContextCUDAImpl::kMaxGradTextureHeight)
                    return false;
                texelIndex = static_cast<uint32_t>(m_simpleRamps.size() * 2);
                m_simpleGradients.emplace(key, texelIndex);
                m_simpleRamps.push_back({ramp[0], ramp[1]});
            }
            location->row = static_cast<uint16_t>(texelIndex / kGradTextureWidth);
            location->col = static_cast<uint16_t>(texelIndex % kGradTextureWidth);
        }
        else
        {
            std::vector<uint32_t> key(stopCount * 2);
            memcpy(key.data(), stops, stopCount * 4);
            memcpy(key.data() + stopCount, colors, stopCount * 4);
            auto it = m_complexGradients.find(key);
            uint16_t row;
            if (it != m_complexGradients.end())
            {
                row = it->second;
            }
            else
            {
                if (data_height(m_simpleRamps.size(), m_complexRamps.size() + 1) > RenderContextCUDAImpl::kMaxGradTextureHeight)
                    return false;
                row = static_cast<uint16_t>(m_complexRamps.size());
                m_complexGradients.emplace(std::move(key), row);
                m_complexRamps.push_back(ref_rcp(gradient));
            }
            location->row = row; // relative to the first complex row until the layout is known
            location->col = ColorRampLocation::kComplexGradientMarker;
        }
        return true;
    }

    // The GradientSpan instances LogicalFlush::writeResources emits (render_context.cpp:1462-1533).
    void writeGradientSpans(uint32_t complexOffsetY, std::vector<GradientSpan>* spans) const
    {
        constexpr uint32_t kOneTexelFixed = 65536 / kGradTextureWidth;
        // constants.glsl:61-63 (a generated header in the reference's build)
        constexpr uint32_t GRAD_SPAN_FLAG_LEFT_BORDER = 0x80000000u, GRAD_SPAN_FLAG_RIGHT_BORDER = 0x40000000u, GRAD_SPAN_FLAG_COMPLEX_BORDER = 0x20000000u;
        for (size_t i = 0; i < m_simpleRamps.size(); ++i)
        {
            // one empty span with one-texel borders to the left and right
            const uint32_t y = static_cast<uint32_t>(i / kGradTextureWidthInSimpleRamps);
            const uint32_t centerXFixed = static_cast<uint32_t>(((i % kGradTextureWidthInSimpleRamps) * 2 + 1) * kOneTexelFixed);
            GradientSpan span;
            span.set(centerXFixed, centerXFixed, y, GRAD_SPAN_FLAG_LEFT_BORDER | GRAD_SPAN_FLAG_RIGHT_BORDER, m_simpleRamps[i][0], m_simpleRamps[i][1]);
            spans->push_back(span);
        }
        for (size_t i = 0; i < m_complexRamps.size(); ++i)
        {
            const Gradient* gradient = m_complexRamps[i].get();
            const float* stops = gradient->stops();
            const ColorInt* colors = gradient->colors();
            const uint32_t y = static_cast<uint32_t>(i) + complexOffsetY;
            const float m = (kGradTextureWidth - 1.f) * kOneTexelFixed, a = .5f * kOneTexelFixed;
            uint32_t lastXFixed = static_cast<uint32_t>(stops[0] * m + a);
            ColorInt lastColor = colors[0];
            for (size_t k = 1; k < gradient->count(); ++k)
            {
                const uint32_t xFixed = static_cast<uint32_t>(stops[k] * m + a);
                uint32_t flags = GRAD_SPAN_FLAG_COMPLEX_BORDER;
                if (k == 1)
                    flags |= GRAD_SPAN_FLAG_LEFT_BORDER;
                if (k == gradient->count() - 1)
                    flags |= GRAD_SPAN_FLAG_RIGHT_BORDER;
                GradientSpan span;
                span.set(lastXFixed, xFixed, y, flags, lastColor, colors[k]);
                spans->push_back(span);
                lastColor = colors[k];
                lastXFixed = xFixed;
            }
        }
    }

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
        size_t clipStackHeight = 0; // clip PATHS: how much of m_clipStack applies to this state
    };
    // RiveRenderer::ClipElement (rive_renderer.hpp:52-72)
    struct ClipElement
    {
        Mat2D matrix;
        rcp<const RiveRenderPath> path;
        uint64_t rawPathMutationID = 0;
        FillRule fillRule = FillRule::nonZero;
        IAABB pixelBounds;
        uint32_t clipID = 0; // assigned every time the element is (re-)rendered to the clip plane
        bool isEquivalent(const Mat2D& matrix_, const RiveRenderPath* path_) const
        {
            return matrix_ == matrix && path_->getRawPathMutationID() == rawPathMutationID && path_->getFillRule() == fillRule;
        }
    };
    constexpr static uint32_t kMaxClipID = 30719; // maxClipID == maxPathID (render_context.cpp:136-139, 484)
    std::vector<ClipElement> m_clipStack;
    uint32_t m_clipContentID = 0; // RenderContext::getClipContentID(): the clip ID now in the clip plane
    uint32_t m_clipCount = 0;
    bool m_hasClipPaths = false;
    std::vector<State> m_stack{State()};
    std::vector<Vec2D> m_points;
    std::vector<uint8_t> m_verbs;
    std::vector<rivecuda_path> m_paths;
    std::vector<rivecuda_clip_rect> m_clipRects;
    // image paints, indexed (1-based) from rivecuda_path::cap >> 8
    std::vector<rivecuda_image_paint> m_imagePaints;
    std::vector<RenderContextCUDAImpl::PlainImageBinding> m_imageBindings;
    std::vector<rcp<Texture>> m_imageTextures; // alive until the flush
    rcp<RiveRenderPath> m_unitRectPath;
    std::vector<RenderContextCUDAImpl::PlainMeshDraw> m_meshDraws; // in draw order
    std::vector<rcp<RenderBuffer>> m_meshBuffers;                  // alive until the flush
    struct GradientDraw
    {
        rcp<const Gradient> gradient;
        ColorRampLocation location;
        Mat2D matrix;
    };
    std::vector<GradientDraw> m_gradientDraws; // indexed (1-based) from rivecuda_path::fill_rule >> 8
    std::unordered_map<uint64_t, uint32_t> m_simpleGradients;         // two colours -> first texel
    std::vector<std::array<ColorInt, 2>> m_simpleRamps;
    std::map<std::vector<uint32_t>, uint16_t> m_complexGradients;     // stops + colours -> row
    std::vector<rcp<const Gradient>> m_complexRamps;
    std::string m_refused;
    bool m_flushFailed = false;
    // delegation
    struct StateCall
    {
        enum Kind
        {
            Transform,
            ClipPath,
            ModulateOpacity
        } kind = Transform;
        Mat2D matrix;
        rcp<RenderPath> path;
        float opacity = 1.f;
    };
    std::vector<std::vector<StateCall>> m_scopes{1}; // one list per open save() scope, parallel to m_stack
    BeginReferenceFrame m_beginReference;
    EndReferenceFrame m_endReference;
    std::unique_ptr<Renderer> m_reference; // open while consecutive draws are being delegated
    size_t m_delegatedDraws = 0;
    bool m_delegateLargeFills = false;
    TriangulationThresholds m_thresholds;
};
} // namespace rive::gpu
