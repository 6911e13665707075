/*
 * TestingWindow backend for RenderContextCUDAImpl, so the reference's GM
 * sources (tests/gm/*.cpp) run unmodified against the CUDA backend -- the
 * analogue of tests/common/testing_window_vulkan_texture.cpp for `--backend
 * cuda`. Also provides the TestingWindow singletons that the
 * reference defines in tests/common/testing_window.cpp (which we cannot
 * compile here because it references every other backend).
 */
#include "testing_window_cuda.hpp"
#include "path_dump.hpp"
#include "cuda_path_renderer.hpp"

#include "rive/renderer/rive_renderer.hpp"

#include <cstdio>
#include <cstdlib>

TestingWindow* s_TestingWindow = nullptr;
TestingWindow::Backend TestingWindow::s_Backend = TestingWindow::Backend::external;
TestingWindow::Target TestingWindow::s_Target = TestingWindow::Target::host;

TestingWindow* TestingWindow::Get()
{
    if (s_TestingWindow == nullptr)
    {
        fprintf(stderr, "TestingWindow::Get(): no window has been Set()\n");
        abort();
    }
    return s_TestingWindow;
}

void TestingWindow::Set(TestingWindow* inWindow) { s_TestingWindow = inWindow; }

void TestingWindow::Destroy()
{
    delete s_TestingWindow;
    s_TestingWindow = nullptr;
}

using namespace rive;
using namespace rive::gpu;

TestingWindowCUDA::TestingWindowCUDA(const RenderContextCUDAImpl::ContextOptions& options) :
    m_renderContext(RenderContextCUDAImpl::MakeContext(options))
{
    if (m_renderContext == nullptr)
    {
        fprintf(stderr, "TestingWindowCUDA: failed to create a CUDA context\n");
        abort();
    }
}

TestingWindowCUDA::~TestingWindowCUDA()
{
    m_renderTarget = nullptr;
    m_renderContext = nullptr;
}

rive::Factory* TestingWindowCUDA::factory() { return m_renderContext.get(); }

void TestingWindowCUDA::resize(int width, int height)
{
    if (m_renderTarget == nullptr || m_width != static_cast<uint32_t>(width) ||
        m_height != static_cast<uint32_t>(height))
    {
        m_renderTarget =
            m_renderContext->static_impl_cast<RenderContextCUDAImpl>()
                ->makeRenderTarget(width, height);
    }
    TestingWindow::resize(width, height);
}

std::unique_ptr<rive::Renderer> TestingWindowCUDA::beginFrame(
    const FrameOptions& options)
{
    RenderContext::FrameDescriptor frameDescriptor = {
        .renderTargetWidth = m_width,
        .renderTargetHeight = m_height,
        .loadAction = options.doClear ? LoadAction::clear
                                      : LoadAction::preserveRenderTarget,
        .clearColor = options.clearColor,
        .msaaSampleCount = 0,
        .disableRasterOrdering = false,
        .triangulationThresholds = options.triangulationThresholds,
        .wireframe = options.wireframe,
        .fillsDisabled = options.fillsDisabled,
        .strokesDisabled = options.strokesDisabled,
        .clockwiseFillOverride = options.clockwiseFillOverride || m_clockwiseFillOverride,
    };
    if (m_hasBudgetOverride)
        frameDescriptor.triangulationThresholds.frameBudgetMs = m_budgetOverride;
    frameDescriptor.virtualTileWidth = m_virtualTileWidth;
    frameDescriptor.virtualTileHeight = m_virtualTileHeight;
    if (m_gpuFrontEnd)
    {
        auto pathRenderer = std::make_unique<CudaPathRenderer>(m_renderContext->static_impl_cast<RenderContextCUDAImpl>(),
                                                               m_renderTarget.get(),
                                                               frameDescriptor.loadAction,
                                                               frameDescriptor.clearColor);
        // Feathers go through the reference's own front end, as flushes of their own in between
        // (CudaPathRenderer::setDelegate); RIVECUDA_FRONT_END_NO_DELEGATE=1 refuses such frames instead.
        if (getenv("RIVECUDA_FRONT_END_NO_DELEGATE") == nullptr)
        {
            pathRenderer->setDelegate(
                [this, frameDescriptor](LoadAction loadAction, ColorInt clearColor) -> std::unique_ptr<rive::Renderer> {
                    RenderContext::FrameDescriptor fd = frameDescriptor;
                    fd.loadAction = loadAction;
                    fd.clearColor = clearColor;
                    m_renderContext->beginFrame(fd);
                    return std::make_unique<RiveRenderer>(m_renderContext.get());
                },
                [this]() { flushPLSContext(nullptr); });
            if (m_delegateLargeFills)
                pathRenderer->setLargeFillDelegation(frameDescriptor.triangulationThresholds);
        }
        m_pathRenderer = pathRenderer.get();
        return pathRenderer;
    }
    m_renderContext->beginFrame(frameDescriptor);
    std::unique_ptr<rive::Renderer> renderer = std::make_unique<RiveRenderer>(m_renderContext.get());
    if (m_pathDump != nullptr && m_pathDump->active)
        renderer = std::make_unique<PathDumpRenderer>(std::move(renderer), m_pathDump);
    return renderer;
}

void TestingWindowCUDA::flushPLSContext(RenderTarget* offscreenRenderTarget)
{
    ++m_frameNumber;
    m_renderContext->flush({
        .renderTarget = offscreenRenderTarget != nullptr ? offscreenRenderTarget
                                                         : m_renderTarget.get(),
        .currentFrameNumber = m_frameNumber,
        .safeFrameNumber = m_frameNumber > 2 ? m_frameNumber - 2 : 0,
    });
}

void TestingWindowCUDA::endFrame(std::vector<uint8_t>* pixelData)
{
    if (m_gpuFrontEnd)
    {
        m_lastFrameRefused = false;
        if (m_pathRenderer == nullptr || !m_pathRenderer->flush())
        {
            fprintf(stderr, "TestingWindowCUDA: --gpu-front-end cannot draw this frame\n");
            if (!m_softRefusal)
                abort();
            m_lastFrameRefused = true;
            m_pathRenderer = nullptr;
            return;
        }
        m_pathRenderer = nullptr;
        if (pixelData != nullptr)
            m_renderTarget->readPixels(pixelData);
        return;
    }
    flushPLSContext(nullptr);
    // Band sharding: one NCCL exchange composites the ranks' bands in rank 0's target.
    m_renderContext->static_impl_cast<RenderContextCUDAImpl>()->gatherBands(m_renderTarget.get(), 0);
    if (m_pathDump != nullptr)
        m_pathDump->active = false; // only the first frame is dumped
    if (pixelData != nullptr)
    {
        m_renderTarget->readPixels(pixelData);
    }
}

rive::gpu::RenderContext* TestingWindowCUDA::renderContext() const
{
    return m_renderContext.get();
}

rive::gpu::RenderTarget* TestingWindowCUDA::renderTarget() const
{
    return m_renderTarget.get();
}
