/*
 * Link-time stubs for the overlay build: the pieces of the reference the headless host does
 * not carry (GLFW input, the assimp importer, the Vulkan Application).  Not reference code.
 */
#include "Application.h"
#include "Core/Core.h"
#include "Core/Input.h"
#include "SceneImporter.h"

namespace PathTracing
{

static Config s_Config = [] {
    Config config;
    config.AssetDirectoryPath = "assets";
    return config;
}();

const Config &Application::GetConfig() { return s_Config; }
void Application::ResetBackgroundTask(BackgroundTaskType) {}
void Application::AddBackgroundTask(BackgroundTaskType, uint32_t) {}
void Application::IncrementBackgroundTaskDone(BackgroundTaskType, uint32_t) {}
void Application::SetBackgroundTaskDone(BackgroundTaskType) {}

GLFWwindow *Input::s_Window = nullptr;
void Input::SetWindow(GLFWwindow *) {}
void Input::LockCursor() {}
void Input::UnlockCursor() {}
bool Input::IsKeyPressed(Key) { return false; }
bool Input::IsMouseButtonPressed(MouseButton) { return false; }
glm::vec2 Input::GetMousePosition() { return glm::vec2(0.0f); }

#ifndef PT_OVERLAY_HAS_ASSIMP
void SceneImporter::Init() {}
void SceneImporter::Shutdown() {}
SceneBuilder &SceneImporter::AddFile(SceneBuilder &, const std::filesystem::path &path, TextureMapping)
{
    throw error("SceneImporter (assimp) is not part of the overlay build: " + path.string());
}
#endif

}
