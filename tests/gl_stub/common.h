/* Stand-in for the reference's src/common.h in the CPU compile test of the PBF_WITH_GL shim branch
 * (tests/test_gl_dropin.py): the GL / glm declarations the reference's header would bring in (glcorew.h, glm.hpp), as
 * stubs that record what was called.  Only what the shim headers and tests/gl_dropin.cpp use. */
#ifndef COMMON_H
#define COMMON_H
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

typedef unsigned int GLuint;
typedef int GLint;
typedef unsigned int GLenum;
typedef ptrdiff_t GLsizeiptr;
typedef ptrdiff_t GLintptr;
#define GL_SHADER_STORAGE_BUFFER 0x90D2
#define GL_COPY_READ_BUFFER 0x8F36
#define GL_COPY_WRITE_BUFFER 0x8F37
#define GL_DYNAMIC_COPY 0x88EA
#define GL_STREAM_COPY 0x88E2
#define GL_DYNAMIC_READ 0x88E9
#define GL_READ_WRITE 0x88BA
#define GL_R32UI 0x8236
#define GL_R8UI 0x8232
#define GL_RED_INTEGER 0x8D94
#define GL_UNSIGNED_INT 0x1405

namespace gl_stub {
struct State {
    GLuint next = 1;
    std::map<GLuint, std::vector<unsigned char>> buffers;
    std::map<GLenum, GLuint> bound;
    int calls = 0;
};
inline State &state() { static State s; return s; }
}  // namespace gl_stub

inline void glGenBuffers(int n, GLuint *out) { for (int i = 0; i < n; i++) { out[i] = gl_stub::state().next++; gl_stub::state().buffers[out[i]]; } gl_stub::state().calls++; }
inline void glDeleteBuffers(int n, const GLuint *b) { for (int i = 0; i < n; i++) gl_stub::state().buffers.erase(b[i]); gl_stub::state().calls++; }
inline void glBindBuffer(GLenum target, GLuint b) { gl_stub::state().bound[target] = b; gl_stub::state().calls++; }
inline void glBufferData(GLenum target, GLsizeiptr size, const void *data, GLenum) {
    auto &v = gl_stub::state().buffers[gl_stub::state().bound[target]];
    v.assign((size_t)size, 0);
    if (data) memcpy(v.data(), data, (size_t)size);
    gl_stub::state().calls++;
}
inline void glClearBufferData(GLenum target, GLenum, GLenum, GLenum, const void *) {
    auto &v = gl_stub::state().buffers[gl_stub::state().bound[target]];
    std::fill(v.begin(), v.end(), 0);
    gl_stub::state().calls++;
}
inline void glCopyBufferSubData(GLenum rt, GLenum wt, GLintptr ro, GLintptr wo, GLsizeiptr size) {
    auto &r = gl_stub::state().buffers[gl_stub::state().bound[rt]];
    auto &w = gl_stub::state().buffers[gl_stub::state().bound[wt]];
    memcpy(w.data() + wo, r.data() + ro, (size_t)size);
    gl_stub::state().calls++;
}
inline void *glMapBuffer(GLenum target, GLenum) { gl_stub::state().calls++; return gl_stub::state().buffers[gl_stub::state().bound[target]].data(); }
inline unsigned char glUnmapBuffer(GLenum) { gl_stub::state().calls++; return 1; }

namespace glm {
struct ivec3 {
    int x, y, z;
    ivec3(int x_ = 0, int y_ = 0, int z_ = 0) : x(x_), y(y_), z(z_) {}
};
}  // namespace glm
#endif /* COMMON_H */
