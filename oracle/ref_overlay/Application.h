/*
 * Stand-in for Path-Tracing/Application.h used ONLY by the overlay build (build.sh): the
 * reference's Scene / SceneManager / ExampleScenes / TextureImporter sources include
 * "Application.h" for Application::GetConfig and the background-task counters; the real header
 * drags in Vulkan and GLFW.  This one declares exactly the members those sources use
 * (Path-Tracing/Application.h:15-36, 41-60) and nothing else.  Not reference code.
 */
#pragma once

#include <atomic>
#include <cstdint>

#include "Core/Config.h"

namespace PathTracing
{

enum class BackgroundTaskType : uint8_t
{
    ShaderCompilation,
    TextureUpload,
    SceneImport,
    Rendering,
};

class Application
{
public:
    static const Config &GetConfig();

    static void ResetBackgroundTask(BackgroundTaskType type);
    static void AddBackgroundTask(BackgroundTaskType type, uint32_t totalCount);
    static void IncrementBackgroundTaskDone(BackgroundTaskType type, uint32_t value = 1);
    static void SetBackgroundTaskDone(BackgroundTaskType type);
};

}
