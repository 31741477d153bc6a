// loader.h -- OBJ mesh loader.  Drop-in for the reference's include/loader.h:12-13.
#pragma once
#include "scene.h"
#include <vector>

// Appends one Object per `f a b c` face of `filename` to `objects`.  Vertices are
// stored as v*scale+offset.  Faces using `/` syntax, more than three indices or
// out-of-range indices are skipped; a missing file prints a message and returns.
void load_obj(const char* filename, std::vector<Object>& objects,
              Vec offset, float scale, Vec albedo, float metallic, float roughness);
