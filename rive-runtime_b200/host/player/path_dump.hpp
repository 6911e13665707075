/*
 * --dump-paths: a Renderer that forwards to the real one and records every plain draw it
 * sees -- solid-colour, src-over, non-feathered nonZero / evenOdd / clockwise fills and strokes: RawPath
 * verbs + points, view matrix, fill rule, colour, stroke thickness / join / cap. That is the
 * input of the GPU path front end (rivecuda_front_end_paths), so its output can be compared
 * with what the reference front end emitted for the very same draws. Anything else (clips,
 * gradients, images, feathers, blend modes) marks the dump as incomplete.
 *
 * File: u32 magic "RPT2", u32 pathCount, u32 complete, u32 0, then per path:
 *   float m[6]; u32 fillRule; u32 color; u32 nVerbs; u32 nPts;
 *   u32 isStroke; float thickness; u32 join; u32 cap;
 *   u8 verbs[nVerbs] (padded to 4); float pts[nPts][2]
 */
#pragma once

#include "rive/renderer.hpp"
#include "rive/math/mat2d.hpp"
#include "rive_render_paint.hpp"
#include "rive_render_path.hpp"

#include <cstdint>
#include <memory>
#include <vector>

struct PathDumpSink
{
    std::vector<uint8_t> blob;
    uint32_t pathCount = 0;
    bool complete = true;
    bool active = true; // cleared after the first frame

    std::vector<uint8_t> file() const
    {
        std::vector<uint8_t> out;
        const uint32_t header[4] = {0x32545052u /* "RPT2" */, pathCount, complete ? 1u : 0u, 0u};
        out.insert(out.end(), reinterpret_cast<const uint8_t*>(header), reinterpret_cast<const uint8_t*>(header) + 16);
        out.insert(out.end(), blob.begin(), blob.end());
        return out;
    }
};

class PathDumpRenderer : public rive::Renderer
{
public:
    PathDumpRenderer(std::unique_ptr<rive::Renderer> inner, PathDumpSink* sink) : m_inner(std::move(inner)), m_sink(sink) {}

    void save() override
    {
        m_stack.push_back(m_stack.back());
        m_inner->save();
    }
    void restore() override
    {
        if (m_stack.size() > 1)
            m_stack.pop_back();
        m_inner->restore();
    }
    void transform(const rive::Mat2D& m) override
    {
        m_stack.back() = m_stack.back() * m;
        m_inner->transform(m);
    }
    void drawPath(rive::RenderPath* path, rive::RenderPaint* paint) override
    {
        using namespace rive;
        auto* rp = static_cast<RiveRenderPath*>(path);
        auto* pt = static_cast<RiveRenderPaint*>(paint);
        const bool plain = pt->getFeather() == 0 && pt->getType() == gpu::PaintType::solidColor &&
                           pt->getBlendMode() == BlendMode::srcOver && pt->getImageTexture() == nullptr;
        // RiveRenderer::drawPath drops these before they reach the front end (rive_renderer.cpp:127-145).
        const bool dropped = rp->getRawPath().empty() || (pt->getIsStroked() && !(pt->getThickness() > 0));
        if (dropped)
        {
        }
        else if (m_sink->active && plain)
        {
            const RawPath& raw = rp->getRawPath();
            const Mat2D& m = m_stack.back();
            for (int i = 0; i < 6; ++i)
                put(m[i]);
            put(static_cast<uint32_t>(rp->getFillRule() == FillRule::evenOdd ? 1 : rp->getFillRule() == FillRule::clockwise ? 2 : 0));
            put(static_cast<uint32_t>(pt->getColor()));
            put(static_cast<uint32_t>(raw.verbs().size()));
            put(static_cast<uint32_t>(raw.points().size()));
            put(static_cast<uint32_t>(pt->getIsStroked() ? 1 : 0));
            put(static_cast<float>(pt->getThickness()));
            put(static_cast<uint32_t>(pt->getJoin()));
            put(static_cast<uint32_t>(pt->getCap()));
            for (PathVerb v : raw.verbs())
                m_sink->blob.push_back(static_cast<uint8_t>(v));
            while (m_sink->blob.size() % 4 != 0)
                m_sink->blob.push_back(0);
            for (Vec2D p : raw.points())
            {
                put(p.x);
                put(p.y);
            }
            ++m_sink->pathCount;
        }
        else if (m_sink->active)
        {
            m_sink->complete = false;
        }
        m_inner->drawPath(path, paint);
    }
    void clipPath(rive::RenderPath* path) override
    {
        incomplete();
        m_inner->clipPath(path);
    }
    void drawImage(const rive::RenderImage* image, rive::ImageSampler sampler, rive::BlendMode blend, float opacity) override
    {
        incomplete();
        m_inner->drawImage(image, sampler, blend, opacity);
    }
    void drawImageMesh(const rive::RenderImage* image,
                       rive::ImageSampler sampler,
                       rive::rcp<rive::RenderBuffer> vertices,
                       rive::rcp<rive::RenderBuffer> uvCoords,
                       rive::rcp<rive::RenderBuffer> indices,
                       uint32_t vertexCount,
                       uint32_t indexCount,
                       rive::BlendMode blend,
                       float opacity) override
    {
        incomplete();
        m_inner->drawImageMesh(image, sampler, vertices, uvCoords, indices, vertexCount, indexCount, blend, opacity);
    }
    void modulateOpacity(float opacity) override
    {
        incomplete();
        m_inner->modulateOpacity(opacity);
    }

private:
    template <typename T> void put(const T& v)
    {
        const uint8_t* b = reinterpret_cast<const uint8_t*>(&v);
        m_sink->blob.insert(m_sink->blob.end(), b, b + sizeof(T));
    }
    void incomplete()
    {
        if (m_sink->active)
            m_sink->complete = false;
    }

    std::unique_ptr<rive::Renderer> m_inner;
    PathDumpSink* m_sink;
    std::vector<rive::Mat2D> m_stack{rive::Mat2D()};
};
