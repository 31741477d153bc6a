// camera.cpp -- CameraController (reference src/camera.cpp:19-163) without SDL.
// The orientation and get_params arithmetic follows the reference (float sin/cos of
// float radians; right = dir x world_up; cx = right*0.5135*aspect, cy = up*0.5135);
// keyboard polling (:85-134) needs a window system and is not part of the hot path.
#include "camera.h"
#include <cmath>

namespace {
inline float to_radians(float deg) { return deg * (M_PI / 180.0f); }
}

CameraController::CameraController(Vec position, Vec look_at) : pos(position) {
    (void)look_at;  // the reference ignores it too (src/camera.cpp:19-24)
    update_camera_vectors();
}

void CameraController::update_camera_vectors() {
    Vec f;
    f.x = std::cos(to_radians(yaw)) * std::cos(to_radians(pitch));
    f.y = std::sin(to_radians(pitch));
    f.z = std::sin(to_radians(yaw)) * std::cos(to_radians(pitch));
    dir = f.norm();
    const Vec world_up{0.f, 1.f, 0.f};
    right = dir.cross(world_up).norm();
    up = right.cross(dir).norm();
}

bool CameraController::process_mouse(float xrel, float yrel) {
    yaw += xrel * mouse_sensitivity;
    pitch -= yrel * mouse_sensitivity;
    if (pitch > 89.0f) pitch = 89.0f;
    if (pitch < -89.0f) pitch = -89.0f;
    update_camera_vectors();
    return true;
}

bool CameraController::update(float) { return false; }

void CameraController::set_angles(float yaw_deg, float pitch_deg) {
    yaw = yaw_deg;
    pitch = pitch_deg;
    if (pitch > 89.0f) pitch = 89.0f;
    if (pitch < -89.0f) pitch = -89.0f;
    update_camera_vectors();
}

CameraParams CameraController::get_params(int width, int height) {
    const float fov_scale = 0.5135f;
    const float aspect = (float)width / height;
    CameraParams p;
    p.pos = pos;
    p.cx = right * (fov_scale * aspect);
    p.cy = up * fov_scale;
    p.dir = dir;
    p.lens_radius = aperture * 0.5f;
    p.focus_dist = focus_dist;
    return p;
}
