// camera.h -- yaw/pitch camera.  Drop-in for the reference's include/camera.h
// (:19 ctor, :31 update, :37 process_mouse, :45 get_params).  Keyboard polling
// is SDL-backed in the reference (src/camera.cpp:85-134); on a headless box
// update() takes explicit key flags instead and update(float) is a no-op.
#pragma once
#include "common.h"
#include "scene.h"

class CameraController {
public:
    CameraController(Vec position, Vec look_at);

    bool update(float delta_time);            // no keyboard on a headless node: returns false
    bool process_mouse(float xrel, float yrel);
    CameraParams get_params(int width, int height);

    float get_aperture() const { return aperture; }
    float get_focus_dist() const { return focus_dist; }

    // headless extensions
    void set_angles(float yaw_deg, float pitch_deg);
    void set_lens(float aperture_, float focus_dist_) { aperture = aperture_; focus_dist = focus_dist_; }

private:
    void update_camera_vectors();
    Vec pos, dir, right, up;
    float yaw = -90.0f;
    float pitch = 0.0f;
    float move_speed = 2.5f;
    float mouse_sensitivity = 0.1f;
    float aperture = 0.0f;
    float focus_dist = 240.0f;
};
