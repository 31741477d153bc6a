// pipeline.h -- display pipeline.  Drop-in for the reference's include/pipeline.h
// (:11-33 Pipeline, :38-48 API).  The worker thread copies the staging buffer
// D2H and tone-maps; the tone-map runs on the device so only 4 B/pixel have to
// cross PCIe when the caller does not need h_accum.
#pragma once
#include <thread>
#include <atomic>
#include <mutex>
#include <condition_variable>
#include <vector>
#include <cstdint>
#include "common.h"

struct Pipeline {
    std::mutex mtx;
    std::condition_variable cv_worker;

    bool quit = false;
    bool worker_busy = false;
    bool frame_ready = false;
    int current_frame = 0;

    Vec* h_accum = nullptr;            // pinned host copy of the accumulation buffer
    Vec* d_staging = nullptr;          // device snapshot the worker reads
    uint32_t* pixel_buffer = nullptr;  // ARGB8888 output
    int width = 0;
    int height = 0;
    size_t size_bytes = 0;

    std::thread worker_thread;
};

void pipeline_init(Pipeline* pipe, Vec* h_accum, Vec* d_staging, uint32_t* pixel_buffer, int w, int h);
bool pipeline_try_dispatch(Pipeline* pipe, int current_gpu_frame);
bool pipeline_check_frame_ready(Pipeline* pipe);
void pipeline_destroy(Pipeline* pipe);
