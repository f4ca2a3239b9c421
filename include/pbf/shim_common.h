/* shim_common.h -- what the shim classes need from the reference's src/common.h (GL types, glm::ivec3).
 * Inside the reference tree define PBF_WITH_GL: the reference's own common.h (glcorew + glm) is used and the SPH shim
 * shares its particle buffers with the renderer through CUDA-GL interop.  Standalone (headless, as in this repo's
 * tests) minimal stand-ins are defined instead. */
#ifndef PBF_SHIM_COMMON_H
#define PBF_SHIM_COMMON_H

#include <stdexcept>
#include <string>

#include "../pbf_c.h"

#ifdef PBF_WITH_GL
/* The reference's src/common.h (glcorew.h, glm, <vector>, <iostream> ...).  This file must not be called common.h itself:
 * a quote include searches the including file's directory first and would find the shim again.  Inside the reference
 * tree every translation unit that uses SPH has already included src/common.h through src/Simulation.h, whose include
 * guard makes this a no-op; otherwise put the reference's src/ on the include path or point PBF_REF_COMMON_H at it. */
#ifndef PBF_REF_COMMON_H
#define PBF_REF_COMMON_H "common.h"
#endif
#include PBF_REF_COMMON_H
#else
typedef unsigned int GLuint;
namespace glm {
struct ivec3 {
    int x, y, z;
    ivec3(int x_ = 0, int y_ = 0, int z_ = 0) : x(x_), y(y_), z(z_) {}
};
}  // namespace glm
#endif

namespace pbf_shim {
/* the reference throws std::runtime_error from its GL wrappers (src/ShaderProgram.cpp:44-116); so do the shims */
inline void check(int rc, const char *what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + pbf_last_error());
}
}  // namespace pbf_shim

#endif
