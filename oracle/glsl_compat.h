/* glsl_compat.h -- just enough of GLSL 4.30 in C++ to compile the reference's compute shaders VERBATIM with g++.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README in pbf_oracle.c's header): used by oracle/ref_harness.cpp, which includes
 * the shader sources of /root/reference/shaders/{sph,radixsort,neighbourcellfinder} after a purely lexical translation
 * (oracle/ref_translate.py: interface blocks -> pointers, `.xyz` -> `.xyz()`, float literals get an `f`, `main` is
 * renamed).  The arithmetic of the shader bodies is therefore the reference's own source text; this header only supplies
 * the vector types, built-in functions and the buffer/texture/image plumbing they are written against.
 *
 * Semantics chosen where GLSL leaves room:
 *  - float arithmetic is IEEE binary32, evaluated in source order (compile with -ffp-contract=off); GLSL allows a driver
 *    to contract and reorder, so this is ONE legal evaluation of the shader -- the same one the C oracle restates;
 *  - dot() sums left to right; length() = sqrtf(dot(v, v)); distance(a, b) = length(a - b);
 *  - dot() of integer vectors converts to float first (GLSL has no integer dot; counting.glsl:56 relies on it);
 *  - texelFetch outside a texture returns `oob_fetch` in every component (GL: undefined without robust access, 0 with
 *    it); imageStore outside an image is dropped (GL: no effect);
 *  - barrier() yields to the work-group scheduler of the harness (ucontext fibers); memory barriers are no-ops because
 *    one work group runs at a time.
 */
#ifndef PBF_GLSL_COMPAT_H
#define PBF_GLSL_COMPAT_H
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef unsigned int uint;

struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    const vec3 &xyz() const { return *this; }
    vec3 &operator+=(const vec3 &o) { x += o.x; y += o.y; z += o.z; return *this; }
    vec3 &operator-=(const vec3 &o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    vec3 &operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
    vec3 &operator/=(float s) { x /= s; y /= s; z /= s; return *this; }
};
inline vec3 operator+(const vec3 &a, const vec3 &b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3 &a, const vec3 &b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3 &a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3 &a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3 &a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }

struct vec4 {
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(const vec3 &v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    vec3 xyz() const { return vec3(x, y, z); }
};

struct ivec3 {
    int x, y, z;
    ivec3() : x(0), y(0), z(0) {}
    ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
    explicit ivec3(const vec3 &v) : x((int)v.x), y((int)v.y), z((int)v.z) {}   /* float -> int truncates */
    int operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline ivec3 operator+(const ivec3 &a, const ivec3 &b) { return ivec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline ivec3 operator*(int s, const ivec3 &a) { return ivec3(s * a.x, s * a.y, s * a.z); }
inline bool operator!=(const ivec3 &a, const ivec3 &b) { return a.x != b.x || a.y != b.y || a.z != b.z; }

struct ivec4 {
    int x, y, z, w;
    ivec4() : x(0), y(0), z(0), w(0) {}
    ivec4(int a, int b, int c, int d) : x(a), y(b), z(c), w(d) {}
    ivec3 xyz() const { return ivec3(x, y, z); }
};

struct bvec4 {
    bool x, y, z, w;
};

struct uvec3 {
    uint x, y, z;
};

struct uvec4 {
    uint x, y, z, w;
    uvec4() : x(0), y(0), z(0), w(0) {}
    uvec4(uint a, uint b, uint c, uint d) : x(a), y(b), z(c), w(d) {}
    explicit uvec4(const bvec4 &b) : x(b.x), y(b.y), z(b.z), w(b.w) {}
    uvec4 &operator+=(const uvec4 &o) { x += o.x; y += o.y; z += o.z; w += o.w; return *this; }
    uint &operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    uint operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    uint &operator[](uint i) { return (*this)[(int)i]; }
    uint operator[](uint i) const { return (*this)[(int)i]; }
};
inline uvec4 operator*(uint s, const uvec4 &a) { return uvec4(s * a.x, s * a.y, s * a.z, s * a.w); }
inline bvec4 equal(const uvec4 &a, const uvec4 &b) { return bvec4{a.x == b.x, a.y == b.y, a.z == b.z, a.w == b.w}; }

inline float dot(const vec3 &a, const vec3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(const ivec3 &a, const ivec3 &b) { return dot(vec3((float)a.x, (float)a.y, (float)a.z), vec3((float)b.x, (float)b.y, (float)b.z)); }
inline float length(const vec3 &v) { return sqrtf(dot(v, v)); }
inline float distance(const vec3 &a, const vec3 &b) { return length(a - b); }
inline vec3 cross(const vec3 &a, const vec3 &b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline float glsl_clamp1(float v, float lo, float hi) { v = v < lo ? lo : v; return v > hi ? hi : v; }   /* min(max(x, lo), hi) */
inline vec3 clamp(const vec3 &v, const vec3 &lo, const vec3 &hi) {
    return vec3(glsl_clamp1(v.x, lo.x, hi.x), glsl_clamp1(v.y, lo.y, hi.y), glsl_clamp1(v.z, lo.z, hi.z));
}
inline int floatBitsToInt(float f) { int i; memcpy(&i, &f, 4); return i; }
inline uint bitfieldExtract(uint value, int offset, int bits) { return (value >> offset) & ((1u << bits) - 1u); }

/* ---- textures, images, buffers: what glBindBufferBase / glBindTexture / glBindImageTexture point at ------------------ */
struct GlUnit {            /* one texture or image unit */
    void *data;
    int comps;             /* components per texel: GL_R32F/R32I/R32UI = 1, GL_RGBA32F/RGBA32I = 4 */
    long n;                /* texels (buffer textures) */
    int dim[3];            /* 3-D textures */
};
struct GlState {
    void *ssbo[8];
    GlUnit tex[8];
    GlUnit img[8];
    int oob_fetch;         /* value of an out-of-range texelFetch */
};
extern GlState gl_state;

struct samplerBuffer { const GlUnit *u; };
struct isamplerBuffer { const GlUnit *u; };
struct isampler3D { const GlUnit *u; };
struct sampler3D { const GlUnit *u; };
struct imageBuffer { const GlUnit *u; };
struct uimageBuffer { const GlUnit *u; };
struct iimageBuffer { const GlUnit *u; };
struct iimage3D { const GlUnit *u; };

inline vec4 texelFetch(const samplerBuffer &s, int i) {
    const GlUnit &u = *s.u;
    if (i < 0 || i >= u.n) { const float o = (float)gl_state.oob_fetch; return vec4(o, o, o, o); }
    const float *p = (const float *)u.data + (long)i * u.comps;
    return u.comps == 4 ? vec4(p[0], p[1], p[2], p[3]) : vec4(p[0], 0.0f, 0.0f, 1.0f);
}
inline ivec4 texelFetch(const isamplerBuffer &s, int i) {
    const GlUnit &u = *s.u;
    if (i < 0 || i >= u.n) { const int o = gl_state.oob_fetch; return ivec4(o, o, o, o); }
    const int *p = (const int *)u.data + (long)i * u.comps;
    return u.comps == 4 ? ivec4(p[0], p[1], p[2], p[3]) : ivec4(p[0], 0, 0, 1);
}
inline ivec4 texelFetch(const isampler3D &s, const ivec3 &c, int /*lod*/) {
    const GlUnit &u = *s.u;
    if (c.x < 0 || c.y < 0 || c.z < 0 || c.x >= u.dim[0] || c.y >= u.dim[1] || c.z >= u.dim[2]) {
        const int o = gl_state.oob_fetch;
        return ivec4(o, o, o, o);
    }
    const int *p = (const int *)u.data + ((long)c.z * u.dim[1] + c.y) * u.dim[0] + c.x;
    return ivec4(p[0], 0, 0, 1);
}
inline vec4 imageLoad(const imageBuffer &s, int i) {
    const GlUnit &u = *s.u;
    if (i < 0 || i >= u.n) return vec4(0, 0, 0, 0);
    const float *p = (const float *)u.data + (long)i * u.comps;
    return u.comps == 4 ? vec4(p[0], p[1], p[2], p[3]) : vec4(p[0], 0.0f, 0.0f, 1.0f);
}
inline uvec4 imageLoad(const uimageBuffer &s, int i) {
    const GlUnit &u = *s.u;
    if (i < 0 || i >= u.n) return uvec4(0, 0, 0, 0);
    return uvec4(((const uint *)u.data)[i], 0, 0, 1);
}
inline void imageStore(const imageBuffer &s, int i, const vec4 &v) {
    const GlUnit &u = *s.u;
    if (i < 0 || i >= u.n) return;
    float *p = (float *)u.data + (long)i * u.comps;
    p[0] = v.x;
    if (u.comps == 4) { p[1] = v.y; p[2] = v.z; p[3] = v.w; }
}
inline void imageStore(const iimageBuffer &s, int i, const ivec4 &v) {
    const GlUnit &u = *s.u;
    if (i < 0 || i >= u.n) return;
    int *p = (int *)u.data + (long)i * u.comps;
    p[0] = v.x;
    if (u.comps == 4) { p[1] = v.y; p[2] = v.z; p[3] = v.w; }
}
inline void imageStore(const iimage3D &s, const ivec3 &c, const ivec4 &v) {
    const GlUnit &u = *s.u;
    if (c.x < 0 || c.y < 0 || c.z < 0 || c.x >= u.dim[0] || c.y >= u.dim[1] || c.z >= u.dim[2]) return;
    ((int *)u.data)[((long)c.z * u.dim[1] + c.y) * u.dim[0] + c.x] = v.x;
}
inline uint imageAtomicOr(const uimageBuffer &s, int i, uint v) {
    const GlUnit &u = *s.u;
    if (i < 0 || i >= u.n) return 0;
    uint *p = (uint *)u.data + i, old = *p;
    *p = old | v;
    return old;
}
inline uint imageAtomicAnd(const uimageBuffer &s, int i, uint v) {
    const GlUnit &u = *s.u;
    if (i < 0 || i >= u.n) return 0;
    uint *p = (uint *)u.data + i, old = *p;
    *p = old & v;
    return old;
}

/* ---- invocation built-ins and work-group barrier (ref_harness.cpp) ----------------------------------------------------- */
extern uvec3 gl_GlobalInvocationID, gl_WorkGroupID;
extern uint gl_LocalInvocationIndex;
void barrier(void);
inline void memoryBarrier(void) {}
inline void memoryBarrierShared(void) {}

#endif
