/*
 * TestingWindow backend for RenderContextCUDAImpl ("--backend cuda").
 * Mirrors tests/common/testing_window_null.cpp / _vulkan_texture.cpp.
 */
#pragma once

#include "common/testing_window.hpp"
#include "render_context_cuda_impl.hpp"

namespace rive::gpu
{
class CudaPathRenderer;
}

class TestingWindowCUDA : public TestingWindow
{
public:
    explicit TestingWindowCUDA(
        const rive::gpu::RenderContextCUDAImpl::ContextOptions& options = {});
    ~TestingWindowCUDA() override;

    rive::Factory* factory() override;
    void resize(int width, int height) override;
    std::unique_ptr<rive::Renderer> beginFrame(const FrameOptions&) override;
    void endFrame(std::vector<uint8_t>* pixelData = nullptr) override;
    void flushPLSContext(
        rive::gpu::RenderTarget* offscreenRenderTarget = nullptr) override;

    rive::gpu::RenderContext* renderContext() const override;
    rive::gpu::RenderTarget* renderTarget() const override;

    // --dump-paths: renderers handed out by beginFrame() record the plain draws of the first
    // frame into `sink` (path_dump.hpp).
    void setPathDump(struct PathDumpSink* sink) { m_pathDump = sink; }

    // --gpu-front-end: beginFrame() hands out a CudaPathRenderer; endFrame()
    // aborts if the frame contained anything but plain fills and strokes.
    void setGpuFrontEnd(bool enabled) { m_gpuFrontEnd = enabled; }
    // --budget-ms given on the command line: it also applies to frames whose FrameOptions carry
    // thresholds of their own (the GMs begin their frames themselves).
    void setTriangulationBudgetOverride(float ms)
    {
        m_budgetOverride = ms;
        m_hasBudgetOverride = true;
    }
    // Sweeps: a frame CudaPathRenderer refuses is reported here instead of aborting.
    // --delegate-large-fills: CudaPathRenderer hands the fills the reference would triangulate to the
    // reference front end too (CudaPathRenderer::setLargeFillDelegation).
    void setDelegateLargeFills(bool enabled) { m_delegateLargeFills = enabled; }
    void setSoftRefusal(bool enabled) { m_softRefusal = enabled; }
    bool lastFrameRefused() const { return m_lastFrameRefused; }

    // FrameDescriptor options the reference exposes per frame: every fill drawn with
    // the clockwise rule; the frame drawn virtual tile by virtual tile.
    void setClockwiseFillOverride(bool enabled) { m_clockwiseFillOverride = enabled; }
    void setVirtualTiles(uint32_t width, uint32_t height)
    {
        m_virtualTileWidth = width;
        m_virtualTileHeight = height;
    }

private:
    struct PathDumpSink* m_pathDump = nullptr;
    bool m_gpuFrontEnd = false;
    bool m_hasBudgetOverride = false;
    float m_budgetOverride = 0;
    bool m_softRefusal = false, m_lastFrameRefused = false;
    bool m_delegateLargeFills = false;
    bool m_clockwiseFillOverride = false;
    uint32_t m_virtualTileWidth = 0, m_virtualTileHeight = 0;
    class rive::gpu::CudaPathRenderer* m_pathRenderer = nullptr; // owned by beginFrame()'s caller
    std::unique_ptr<rive::gpu::RenderContext> m_renderContext;
    rive::rcp<rive::gpu::RenderTargetCUDA> m_renderTarget;
    uint64_t m_frameNumber = 0;
};
