// image_io.h -- PPM snapshot writer.  Drop-in for the reference's include/image_io.h:6.
#pragma once
#include "common.h"

// Tone-maps accum/frame with toInt and writes logs/<timestamp>_Frame..ppm (P6).
void save_snapshot(const Vec* h_accum, int w, int h, int frame, float focus_dist, float aperture);

// P6 reader used for textures (reference src/renderer.cu:36-76).  Returns a
// malloc'ed RGB8 buffer or NULL.
unsigned char* load_ppm(const char* filename, int* w, int* h);
