// puckswarm_b200 -- math primitives of the batched JBox2D-2.1.2-semantics engine (device code).
//
// Operation order of every expression follows JBox2D / Box2D v2.1.2 because float
// rounding is part of the contract: the library is built with --fmad=false (Java never fuses).
// The same headers also compile as plain C++ (PS_HOST_EMU) for the host emulation harness used by the
// CPU-only unit tests in tests/; the product library contains device code only.
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define PS_HD __host__ __device__ __forceinline__
#define PS_HDN __host__ __device__ __noinline__
#else
#define PS_HD inline
#define PS_HDN
#endif

namespace ps {

// ---- org.jbox2d.common.Settings ----
#define PS_PI 3.14159265359f
static constexpr float kPi = PS_PI;
static constexpr float kTwoPi = (float)(3.14159265358979323846 * 2);
static constexpr float kHalfPi = PS_PI / 2;
static constexpr float kEpsilon = 1.1920928955078125e-7f;
static constexpr float kMaxFloat = FLT_MAX;
static constexpr float kAabbExtension = 0.1f;
static constexpr float kAabbMultiplier = 2.0f;
static constexpr float kLinearSlop = 0.005f;
static constexpr float kPolygonRadius = 2.0f * 0.005f;
static constexpr float kVelocityThreshold = 1.0f;
static constexpr float kMaxLinearCorrection = 0.2f;
static constexpr float kMaxTranslation = 2.0f;
static constexpr float kMaxTranslationSquared = 4.0f;
static constexpr float kMaxRotation = 0.5f * PS_PI;
static constexpr float kMaxRotationSquared = (0.5f * PS_PI) * (0.5f * PS_PI);
static constexpr float kContactBaumgarte = 0.2f;
static constexpr float kToiBaumgarte = 0.75f;
static constexpr int kMaxTOIContacts = 32;
static constexpr float kLutPrecision = 0.00011f;
static constexpr int kLutLength = 57120;

struct V2 {
    float x, y;
};
PS_HD V2 mk(float x, float y) {
    V2 r;
    r.x = x;
    r.y = y;
    return r;
}
PS_HD V2 operator+(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }
PS_HD V2 operator-(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }
PS_HD V2 operator-(V2 a) { return mk(-a.x, -a.y); }
PS_HD V2 operator*(float s, V2 a) { return mk(s * a.x, s * a.y); }
PS_HD float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
PS_HD float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
PS_HD V2 crossVS(V2 a, float s) { return mk(s * a.y, -s * a.x); }
PS_HD V2 crossSV(float s, V2 a) { return mk(-s * a.y, s * a.x); }
PS_HD float distSq(V2 a, V2 b) {
    V2 c = a - b;
    return dot(c, c);
}
PS_HD float len(V2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
PS_HD float dist(V2 a, V2 b) { return len(a - b); }
// Vec2.normalize: returns the length, leaves the vector alone below EPSILON
PS_HD float normalize(V2& a) {
    float l = len(a);
    if (l < kEpsilon) return 0.0f;
    float inv = 1.0f / l;
    a.x *= inv;
    a.y *= inv;
    return l;
}
PS_HD float fminf_(float a, float b) { return a < b ? a : b; }
PS_HD float fmaxf_(float a, float b) { return a > b ? a : b; }
// MathUtils.clamp = max(lo, min(a, hi))
PS_HD float clampf(float a, float lo, float hi) { return fmaxf_(lo, fminf_(a, hi)); }

// Transform as (position, cos, sin): R.col1 = (c, s), R.col2 = (-s, c)
struct Xf {
    float px, py, c, s;
};
PS_HD V2 mulR(const Xf& T, V2 v) { return mk(T.c * v.x + (-T.s) * v.y, T.s * v.x + T.c * v.y); }
PS_HD V2 mulTR(const Xf& T, V2 v) { return mk(v.x * T.c + v.y * T.s, v.x * (-T.s) + v.y * T.c); }
PS_HD V2 mulX(const Xf& T, V2 v) { return mk(T.px + T.c * v.x + (-T.s) * v.y, T.py + T.s * v.x + T.c * v.y); }
PS_HD V2 mulTX(const Xf& T, V2 v) { return mulTR(T, mk(v.x - T.px, v.y - T.py)); }

// MathUtils.sin / cos with Settings.SINCOS_LUT_ENABLED: table of (float)Math.sin(i * 0.00011f)
struct Trig {
    const float* lut;
    int useLut;
};
// (int)(x / LUT_PRECISION + 0.5f) for x in [0, 2 pi]. On the device the quotient comes from a multiplication by the
// rounded reciprocal and one fused residual correction instead of the IEEE division sequence: the index is identical for
// every float of the domain (exhaustive check: tools/check_lut_index.c, 1.09e9 values, 0 mismatches).
PS_HD int lutIndex(float x) {
#if defined(__CUDA_ARCH__)
    const float rp = 1.0f / kLutPrecision;
    float q0 = x * rp;
    float r = __fmaf_rn(-q0, kLutPrecision, x);
    float q = __fmaf_rn(r, rp, q0);
    return (int)(q + 0.5f);
#else
    return (int)(x / kLutPrecision + 0.5f);
#endif
}
PS_HD float psSin(const Trig& t, float x) {
    if (t.useLut) {
        // x %= TWOPI; fmod is exact, and the identity below 2 pi (the common case) -- skip the slow path there
        if (!(fabsf(x) < kTwoPi)) x = fmodf(x, kTwoPi);
        if (x < 0) x += kTwoPi;
        int idx = lutIndex(x);
        // idx % LUT_LENGTH: after the reduction x is in [0, 2 pi], so idx <= LUT_LENGTH and the modulo is one subtraction
        if (idx >= kLutLength) idx -= kLutLength;
        return t.lut[idx];
    }
    return (float)sin((double)x);
}
PS_HD float psCos(const Trig& t, float x) {
    if (t.useLut) return psSin(t, kHalfPi - x);
    return (float)cos((double)x);
}
// Body.synchronizeTransform
PS_HD Xf makeXf(const Trig& t, float cx, float cy, float a, float lcx, float lcy) {
    Xf T;
    T.c = psCos(t, a);
    T.s = psSin(t, a);
    T.px = cx - (T.c * lcx + (-T.s) * lcy);
    T.py = cy - (T.s * lcx + T.c * lcy);
    return T;
}

PS_HD int ctz32(uint32_t v) {   // v != 0
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}

struct Aabb {
    float lx, ly, ux, uy;
};
PS_HD bool aabbOverlap(const Aabb& a, const Aabb& b) {
    float d1x = b.lx - a.ux, d1y = b.ly - a.uy;
    float d2x = a.lx - b.ux, d2y = a.ly - b.uy;
    if (d1x > 0.0f || d1y > 0.0f) return false;
    if (d2x > 0.0f || d2y > 0.0f) return false;
    return true;
}
PS_HD bool aabbContains(const Aabb& a, const Aabb& o) { return a.lx <= o.lx && a.ly <= o.ly && o.ux <= a.ux && o.uy <= a.uy; }

}  // namespace ps
