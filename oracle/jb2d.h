// ORACLE (test infrastructure only -- never linked into or called by the product path).
//
// CPU restatement of the part of org.jbox2d:jbox2d-library:2.1.2.2 (pinned in /root/reference/ivy.xml:10;
// the jar/source is NOT in the reference tree) that PuckSwarm's World.step exercises:
// dynamic bodies made of circle / polygon fixtures, one static wall body, contact cache with fat-AABB
// broad-phase, circle-circle / polygon-circle / polygon-polygon manifolds, island DFS, warm-started
// sequential-impulse velocity solve, mass-scaled position solve and TOI against the static body.
// It follows the published Box2D v2.1.2 algorithm (which JBox2D 2.1.2.x ports) as recorded in
// SURVEY.md App. A; PARITY IS UNPINNED (the reference ships no tests or golden vectors for this path,
// and no JVM/JBox2D exists here to generate any). Reference call sites that define how it is used:
// src/arena/RunOffline.java:94,108; src/arena/Robot.java:70-78,95-163,175-182; src/arena/Puck.java:30-55;
// src/arena/DotPuck.java:25-33; src/utils/FixtureUtils.java:22-74.
//
// Every float expression keeps Box2D's operation order; build with -ffp-contract=off (Java never fuses).
#pragma once
#include <cstdint>
#include <cmath>
#include <cfloat>
#include <vector>
#include <algorithm>
#include <cstring>

namespace jb {

// ---- switches for the JBox2D details SURVEY marks uncertain ----
struct Switches {
    bool sincosLut = true;            // Settings.SINCOS_LUT_ENABLED (A.2)
    bool toiEnabled = true;           // World.m_continuousPhysics (A.7)
    bool posSolveMassScaled = true;   // ContactSolver.solvePositionConstraints mass scaling (A.6 step 7)
    bool partitionStaticLast = true;  // Island.solve contact partition (Box2D 2.1.2 b2Island::Solve)
};

// ---- Settings (A.1) ----
static const float kPi = 3.14159265359f;          // Box2D b2_pi; JBox2D MathUtils.PI = (float)Math.PI (same float)
static const float kTwoPi = (float)(M_PI * 2);
static const float kHalfPi = kPi / 2;
static const float kEpsilon = 1.1920928955078125e-7f;
static const float kMaxFloat = FLT_MAX;
static const float kAabbExtension = 0.1f;
static const float kAabbMultiplier = 2.0f;
static const float kLinearSlop = 0.005f;
static const float kPolygonRadius = 2.0f * kLinearSlop;
static const float kVelocityThreshold = 1.0f;
static const float kMaxLinearCorrection = 0.2f;
static const float kMaxTranslation = 2.0f;
static const float kMaxTranslationSquared = kMaxTranslation * kMaxTranslation;
static const float kMaxRotation = 0.5f * kPi;
static const float kMaxRotationSquared = kMaxRotation * kMaxRotation;
static const float kContactBaumgarte = 0.2f;
static const int kMaxTOIContacts = 32;
static const int kMaxManifoldPoints = 2;
static const int kMaxPolygonVertices = 8;

// ---- MathUtils sin/cos (A.2) ----
static const float kLutPrecision = 0.00011f;
static const int kLutLength = 57120;  // (int)ceil(2*pi / 0.00011)
inline const float* sinTable() {
    static std::vector<float> t;
    if (t.empty()) {
        t.resize(kLutLength);
        for (int i = 0; i < kLutLength; i++) t[i] = (float)std::sin((double)(i * kLutPrecision));
    }
    return t.data();
}
inline float sinLUT(float x) {
    x = std::fmod(x, kTwoPi);
    while (x < 0) x += kTwoPi;
    int idx = (int)(x / kLutPrecision + 0.5f);   // MathUtils.round = floor(x + .5f), x >= 0
    return sinTable()[idx % kLutLength];
}
struct Trig {
    bool lut;
    float sin(float x) const { return lut ? sinLUT(x) : (float)std::sin((double)x); }
    float cos(float x) const { return lut ? sinLUT(kHalfPi - x) : (float)std::cos((double)x); }
};

// ---- vectors ----
struct Vec2 {
    float x, y;
    Vec2() : x(0), y(0) {}
    Vec2(float x_, float y_) : x(x_), y(y_) {}
    void set(float x_, float y_) { x = x_; y = y_; }
    Vec2 operator-() const { return Vec2(-x, -y); }
    void operator+=(const Vec2& v) { x += v.x; y += v.y; }
    void operator-=(const Vec2& v) { x -= v.x; y -= v.y; }
    void operator*=(float a) { x *= a; y *= a; }
    float length() const { return std::sqrt(x * x + y * y); }
    float lengthSquared() const { return x * x + y * y; }
    float normalize() {
        float len = length();
        if (len < kEpsilon) return 0.0f;
        float inv = 1.0f / len;
        x *= inv;
        y *= inv;
        return len;
    }
};
inline Vec2 operator+(const Vec2& a, const Vec2& b) { return Vec2(a.x + b.x, a.y + b.y); }
inline Vec2 operator-(const Vec2& a, const Vec2& b) { return Vec2(a.x - b.x, a.y - b.y); }
inline Vec2 operator*(float s, const Vec2& a) { return Vec2(s * a.x, s * a.y); }
inline float dot(const Vec2& a, const Vec2& b) { return a.x * b.x + a.y * b.y; }
inline float cross(const Vec2& a, const Vec2& b) { return a.x * b.y - a.y * b.x; }
inline Vec2 cross(const Vec2& a, float s) { return Vec2(s * a.y, -s * a.x); }
inline Vec2 cross(float s, const Vec2& a) { return Vec2(-s * a.y, s * a.x); }
inline float distanceSquared(const Vec2& a, const Vec2& b) {
    Vec2 c = a - b;
    return dot(c, c);
}
inline float distance(const Vec2& a, const Vec2& b) { return (a - b).length(); }
// MathUtils.min / max / clamp (JBox2D): a < b ? a : b, a > b ? a : b, max(lo, min(a, hi))
inline float minf(float a, float b) { return a < b ? a : b; }
inline float maxf(float a, float b) { return a > b ? a : b; }
inline float clampf(float a, float lo, float hi) { return maxf(lo, minf(a, hi)); }
inline Vec2 vmin(const Vec2& a, const Vec2& b) { return Vec2(minf(a.x, b.x), minf(a.y, b.y)); }
inline Vec2 vmax(const Vec2& a, const Vec2& b) { return Vec2(maxf(a.x, b.x), maxf(a.y, b.y)); }

struct Mat22 {
    Vec2 col1, col2;
    void set(float angle, const Trig& tr) {
        float c = tr.cos(angle), s = tr.sin(angle);
        col1.x = c;
        col2.x = -s;
        col1.y = s;
        col2.y = c;
    }
    Mat22 getInverse() const {
        float a = col1.x, b = col2.x, c = col1.y, d = col2.y;
        Mat22 B;
        float det = a * d - b * c;
        if (det != 0.0f) det = 1.0f / det;
        B.col1.x = det * d;
        B.col2.x = -det * b;
        B.col1.y = -det * c;
        B.col2.y = det * a;
        return B;
    }
};
inline Vec2 mul(const Mat22& A, const Vec2& v) {
    return Vec2(A.col1.x * v.x + A.col2.x * v.y, A.col1.y * v.x + A.col2.y * v.y);
}
inline Vec2 mulT(const Mat22& A, const Vec2& v) { return Vec2(dot(v, A.col1), dot(v, A.col2)); }
struct Transform {
    Vec2 position;
    Mat22 R;
};
inline Vec2 mul(const Transform& T, const Vec2& v) {
    float x = T.position.x + T.R.col1.x * v.x + T.R.col2.x * v.y;
    float y = T.position.y + T.R.col1.y * v.x + T.R.col2.y * v.y;
    return Vec2(x, y);
}
inline Vec2 mulT(const Transform& T, const Vec2& v) { return mulT(T.R, v - T.position); }

struct Sweep {
    Vec2 localCenter, c0, c;
    float a0 = 0, a = 0;
    void getTransform(Transform* xf, float alpha, const Trig& tr) const {
        xf->position = (1.0f - alpha) * c0 + alpha * c;
        float angle = (1.0f - alpha) * a0 + alpha * a;
        xf->R.set(angle, tr);
        xf->position -= mul(xf->R, localCenter);
    }
    void advance(float t) {
        c0 = (1.0f - t) * c0 + t * c;
        a0 = (1.0f - t) * a0 + t * a;
    }
    void normalize() {
        float d = kTwoPi * std::floor(a0 / kTwoPi);
        a0 -= d;
        a -= d;
    }
};

struct AABB {
    Vec2 lower, upper;
    bool contains(const AABB& o) const {
        bool r = true;
        r = r && lower.x <= o.lower.x;
        r = r && lower.y <= o.lower.y;
        r = r && o.upper.x <= upper.x;
        r = r && o.upper.y <= upper.y;
        return r;
    }
    void combine(const AABB& a, const AABB& b) {
        lower = vmin(a.lower, b.lower);
        upper = vmax(a.upper, b.upper);
    }
};
inline bool testOverlap(const AABB& a, const AABB& b) {
    Vec2 d1 = b.lower - a.upper, d2 = a.lower - b.upper;
    if (d1.x > 0.0f || d1.y > 0.0f) return false;
    if (d2.x > 0.0f || d2.y > 0.0f) return false;
    return true;
}

// ---- shapes (A.9 "Shapes") ----
enum ShapeType { CIRCLE = 0, POLYGON = 1 };
struct MassData {
    float mass = 0, I = 0;
    Vec2 center;
};
struct Shape {
    int type = CIRCLE;
    float radius = 0;
    Vec2 p;  // circle centre
    int count = 0;
    Vec2 v[kMaxPolygonVertices], n[kMaxPolygonVertices];
    Vec2 centroid;

    static Shape circle(float r, Vec2 p_ = Vec2()) {
        Shape s;
        s.type = CIRCLE;
        s.radius = r;
        s.p = p_;
        return s;
    }
    static Vec2 computeCentroid(const Vec2* vs, int count) {
        Vec2 c;
        float area = 0.0f;
        if (count == 2) {
            c = 0.5f * (vs[0] + vs[1]);
            return c;
        }
        Vec2 pRef(0, 0);
        const float inv3 = 1.0f / 3.0f;
        for (int i = 0; i < count; ++i) {
            Vec2 p1 = pRef, p2 = vs[i], p3 = i + 1 < count ? vs[i + 1] : vs[0];
            Vec2 e1 = p2 - p1, e2 = p3 - p1;
            float D = cross(e1, e2);
            float triangleArea = 0.5f * D;
            area += triangleArea;
            c += triangleArea * inv3 * (p1 + p2 + p3);
        }
        c *= 1.0f / area;
        return c;
    }
    static Shape polygon(const Vec2* vs, int count) {
        Shape s;
        s.type = POLYGON;
        s.radius = kPolygonRadius;
        s.count = count;
        for (int i = 0; i < count; i++) s.v[i] = vs[i];
        for (int i = 0; i < count; i++) {
            int i2 = i + 1 < count ? i + 1 : 0;
            Vec2 edge = s.v[i2] - s.v[i];
            s.n[i] = cross(edge, 1.0f);
            s.n[i].normalize();
        }
        s.centroid = computeCentroid(s.v, count);
        return s;
    }
    static Shape edge(Vec2 v1, Vec2 v2) {  // PolygonShape.setAsEdge
        Shape s;
        s.type = POLYGON;
        s.radius = kPolygonRadius;
        s.count = 2;
        s.v[0] = v1;
        s.v[1] = v2;
        s.centroid = 0.5f * (v1 + v2);
        s.n[0] = cross(v2 - v1, 1.0f);
        s.n[0].normalize();
        s.n[1] = -s.n[0];
        return s;
    }
    void computeAABB(AABB* aabb, const Transform& xf) const {
        if (type == CIRCLE) {
            Vec2 c = xf.position + mul(xf.R, p);
            aabb->lower.set(c.x - radius, c.y - radius);
            aabb->upper.set(c.x + radius, c.y + radius);
        } else {
            Vec2 lo = mul(xf, v[0]), up = lo;
            for (int i = 1; i < count; ++i) {
                Vec2 w = mul(xf, v[i]);
                lo = vmin(lo, w);
                up = vmax(up, w);
            }
            Vec2 r(radius, radius);
            aabb->lower = lo - r;
            aabb->upper = up + r;
        }
    }
    void computeMass(MassData* md, float density) const {
        if (type == CIRCLE) {
            md->mass = density * kPi * radius * radius;
            md->center = p;
            md->I = md->mass * (0.5f * radius * radius + dot(p, p));
            return;
        }
        if (count == 2) {
            md->center = 0.5f * (v[0] + v[1]);
            md->mass = 0.0f;
            md->I = 0.0f;
            return;
        }
        Vec2 center(0, 0);
        float area = 0, I = 0;
        Vec2 pRef(0, 0);
        const float k_inv3 = 1.0f / 3.0f;
        for (int i = 0; i < count; ++i) {
            Vec2 p1 = pRef, p2 = v[i], p3 = i + 1 < count ? v[i + 1] : v[0];
            Vec2 e1 = p2 - p1, e2 = p3 - p1;
            float D = cross(e1, e2);
            float triangleArea = 0.5f * D;
            area += triangleArea;
            center += triangleArea * k_inv3 * (p1 + p2 + p3);
            float px = p1.x, py = p1.y, ex1 = e1.x, ey1 = e1.y, ex2 = e2.x, ey2 = e2.y;
            float intx2 = k_inv3 * (0.25f * (ex1 * ex1 + ex2 * ex1 + ex2 * ex2) + (px * ex1 + px * ex2)) + 0.5f * px * px;
            float inty2 = k_inv3 * (0.25f * (ey1 * ey1 + ey2 * ey1 + ey2 * ey2) + (py * ey1 + py * ey2)) + 0.5f * py * py;
            I += D * (intx2 + inty2);
        }
        md->mass = density * area;
        center *= 1.0f / area;
        md->center = center;
        md->I = density * I;
    }
    bool testPoint(const Transform& xf, const Vec2& pt) const {
        if (type == CIRCLE) {
            Vec2 center = xf.position + mul(xf.R, p);
            Vec2 d = pt - center;
            return dot(d, d) <= radius * radius;
        }
        Vec2 pLocal = mulT(xf.R, pt - xf.position);
        for (int i = 0; i < count; ++i) {
            float d = dot(n[i], pLocal - v[i]);
            if (d > 0.0f) return false;
        }
        return true;
    }
};

// ---- manifolds ----
enum ManifoldType { M_CIRCLES = 0, M_FACE_A = 1, M_FACE_B = 2 };
struct ContactID {
    uint8_t referenceEdge = 0, incidentEdge = 0, incidentVertex = 0, flip = 0;
    uint32_t key() const { return referenceEdge | (incidentEdge << 8) | (incidentVertex << 16) | (flip << 24); }
    // 8-bit packing used by the exchange format (include/puckswarm_b200.h contact_info)
    uint8_t pack() const { return (uint8_t)(referenceEdge | (incidentEdge << 3) | (incidentVertex << 6) | (flip << 7)); }
    static ContactID unpack(uint8_t b) {
        ContactID id;
        id.referenceEdge = b & 7;
        id.incidentEdge = (b >> 3) & 7;
        id.incidentVertex = (b >> 6) & 1;
        id.flip = (b >> 7) & 1;
        return id;
    }
};
struct ManifoldPoint {
    Vec2 localPoint;
    float normalImpulse = 0, tangentImpulse = 0;
    ContactID id;
};
struct Manifold {
    ManifoldPoint points[kMaxManifoldPoints];
    Vec2 localNormal, localPoint;
    int type = M_CIRCLES;
    int pointCount = 0;
};
struct WorldManifold {
    Vec2 normal;
    Vec2 points[kMaxManifoldPoints];
    void initialize(const Manifold* m, const Transform& xfA, float radiusA, const Transform& xfB, float radiusB) {
        if (m->pointCount == 0) return;
        switch (m->type) {
            case M_CIRCLES: {
                Vec2 pointA = mul(xfA, m->localPoint);
                Vec2 pointB = mul(xfB, m->points[0].localPoint);
                Vec2 nrm(1.0f, 0.0f);
                if (distanceSquared(pointA, pointB) > kEpsilon * kEpsilon) {
                    nrm = pointB - pointA;
                    nrm.normalize();
                }
                normal = nrm;
                Vec2 cA = pointA + radiusA * nrm;
                Vec2 cB = pointB - radiusB * nrm;
                points[0] = 0.5f * (cA + cB);
            } break;
            case M_FACE_A: {
                Vec2 nrm = mul(xfA.R, m->localNormal);
                Vec2 planePoint = mul(xfA, m->localPoint);
                normal = nrm;
                for (int i = 0; i < m->pointCount; ++i) {
                    Vec2 clipPoint = mul(xfB, m->points[i].localPoint);
                    Vec2 cA = clipPoint + (radiusA - dot(clipPoint - planePoint, nrm)) * nrm;
                    Vec2 cB = clipPoint - radiusB * nrm;
                    points[i] = 0.5f * (cA + cB);
                }
            } break;
            case M_FACE_B: {
                Vec2 nrm = mul(xfB.R, m->localNormal);
                Vec2 planePoint = mul(xfB, m->localPoint);
                for (int i = 0; i < m->pointCount; ++i) {
                    Vec2 clipPoint = mul(xfA, m->points[i].localPoint);
                    Vec2 cB = clipPoint + (radiusB - dot(clipPoint - planePoint, nrm)) * nrm;
                    Vec2 cA = clipPoint - radiusA * nrm;
                    points[i] = 0.5f * (cA + cB);
                }
                normal = -nrm;
            } break;
        }
    }
};

// ---- narrow phase (A.4, A.9) ----
inline void collideCircles(Manifold* m, const Shape* A, const Transform& xfA, const Shape* B, const Transform& xfB) {
    m->pointCount = 0;
    Vec2 pA = mul(xfA, A->p), pB = mul(xfB, B->p);
    Vec2 d = pB - pA;
    float distSqr = dot(d, d);
    float radius = A->radius + B->radius;
    if (distSqr > radius * radius) return;
    m->type = M_CIRCLES;
    m->localPoint = A->p;
    m->localNormal = Vec2();
    m->pointCount = 1;
    m->points[0].localPoint = B->p;
    m->points[0].id = ContactID();
}

inline void collidePolygonAndCircle(Manifold* m, const Shape* A, const Transform& xfA, const Shape* B, const Transform& xfB) {
    m->pointCount = 0;
    Vec2 c = mul(xfB, B->p);
    Vec2 cLocal = mulT(xfA, c);
    int normalIndex = 0;
    float separation = -kMaxFloat;
    float radius = A->radius + B->radius;
    int vertexCount = A->count;
    const Vec2* vertices = A->v;
    const Vec2* normals = A->n;
    for (int i = 0; i < vertexCount; ++i) {
        float s = dot(normals[i], cLocal - vertices[i]);
        if (s > radius) return;
        if (s > separation) {
            separation = s;
            normalIndex = i;
        }
    }
    int vertIndex1 = normalIndex;
    int vertIndex2 = vertIndex1 + 1 < vertexCount ? vertIndex1 + 1 : 0;
    Vec2 v1 = vertices[vertIndex1], v2 = vertices[vertIndex2];
    if (separation < kEpsilon) {
        m->pointCount = 1;
        m->type = M_FACE_A;
        m->localNormal = normals[normalIndex];
        m->localPoint = 0.5f * (v1 + v2);
        m->points[0].localPoint = B->p;
        m->points[0].id = ContactID();
        return;
    }
    float u1 = dot(cLocal - v1, v2 - v1);
    float u2 = dot(cLocal - v2, v1 - v2);
    if (u1 <= 0.0f) {
        if (distanceSquared(cLocal, v1) > radius * radius) return;
        m->pointCount = 1;
        m->type = M_FACE_A;
        m->localNormal = cLocal - v1;
        m->localNormal.normalize();
        m->localPoint = v1;
        m->points[0].localPoint = B->p;
        m->points[0].id = ContactID();
    } else if (u2 <= 0.0f) {
        if (distanceSquared(cLocal, v2) > radius * radius) return;
        m->pointCount = 1;
        m->type = M_FACE_A;
        m->localNormal = cLocal - v2;
        m->localNormal.normalize();
        m->localPoint = v2;
        m->points[0].localPoint = B->p;
        m->points[0].id = ContactID();
    } else {
        Vec2 faceCenter = 0.5f * (v1 + v2);
        float sep = dot(cLocal - faceCenter, normals[vertIndex1]);
        if (sep > radius) return;
        m->pointCount = 1;
        m->type = M_FACE_A;
        m->localNormal = normals[vertIndex1];
        m->localPoint = faceCenter;
        m->points[0].localPoint = B->p;
        m->points[0].id = ContactID();
    }
}

struct ClipVertex {
    Vec2 v;
    ContactID id;
};
inline float edgeSeparation(const Shape* poly1, const Transform& xf1, int edge1, const Shape* poly2, const Transform& xf2) {
    Vec2 normal1World = mul(xf1.R, poly1->n[edge1]);
    Vec2 normal1 = mulT(xf2.R, normal1World);
    int index = 0;
    float minDot = kMaxFloat;
    for (int i = 0; i < poly2->count; ++i) {
        float d = dot(poly2->v[i], normal1);
        if (d < minDot) {
            minDot = d;
            index = i;
        }
    }
    Vec2 v1 = mul(xf1, poly1->v[edge1]);
    Vec2 v2 = mul(xf2, poly2->v[index]);
    return dot(v2 - v1, normal1World);
}
inline float findMaxSeparation(int* edgeIndex, const Shape* poly1, const Transform& xf1, const Shape* poly2, const Transform& xf2) {
    int count1 = poly1->count;
    Vec2 d = mul(xf2, poly2->centroid) - mul(xf1, poly1->centroid);
    Vec2 dLocal1 = mulT(xf1.R, d);
    int edge = 0;
    float maxDot = -kMaxFloat;
    for (int i = 0; i < count1; ++i) {
        float dt = dot(poly1->n[i], dLocal1);
        if (dt > maxDot) {
            maxDot = dt;
            edge = i;
        }
    }
    float s = edgeSeparation(poly1, xf1, edge, poly2, xf2);
    int prevEdge = edge - 1 >= 0 ? edge - 1 : count1 - 1;
    float sPrev = edgeSeparation(poly1, xf1, prevEdge, poly2, xf2);
    int nextEdge = edge + 1 < count1 ? edge + 1 : 0;
    float sNext = edgeSeparation(poly1, xf1, nextEdge, poly2, xf2);
    int bestEdge, increment;
    float bestSeparation;
    if (sPrev > s && sPrev > sNext) {
        increment = -1;
        bestEdge = prevEdge;
        bestSeparation = sPrev;
    } else if (sNext > s) {
        increment = 1;
        bestEdge = nextEdge;
        bestSeparation = sNext;
    } else {
        *edgeIndex = edge;
        return s;
    }
    for (;;) {
        if (increment == -1)
            edge = bestEdge - 1 >= 0 ? bestEdge - 1 : count1 - 1;
        else
            edge = bestEdge + 1 < count1 ? bestEdge + 1 : 0;
        s = edgeSeparation(poly1, xf1, edge, poly2, xf2);
        if (s > bestSeparation) {
            bestEdge = edge;
            bestSeparation = s;
        } else
            break;
    }
    *edgeIndex = bestEdge;
    return bestSeparation;
}
inline void findIncidentEdge(ClipVertex c[2], const Shape* poly1, const Transform& xf1, int edge1, const Shape* poly2, const Transform& xf2) {
    Vec2 normal1 = mulT(xf2.R, mul(xf1.R, poly1->n[edge1]));
    int index = 0;
    float minDot = kMaxFloat;
    for (int i = 0; i < poly2->count; ++i) {
        float d = dot(normal1, poly2->n[i]);
        if (d < minDot) {
            minDot = d;
            index = i;
        }
    }
    int i1 = index, i2 = i1 + 1 < poly2->count ? i1 + 1 : 0;
    c[0].v = mul(xf2, poly2->v[i1]);
    c[0].id.referenceEdge = (uint8_t)edge1;
    c[0].id.incidentEdge = (uint8_t)i1;
    c[0].id.incidentVertex = 0;
    c[0].id.flip = 0;
    c[1].v = mul(xf2, poly2->v[i2]);
    c[1].id.referenceEdge = (uint8_t)edge1;
    c[1].id.incidentEdge = (uint8_t)i2;
    c[1].id.incidentVertex = 1;
    c[1].id.flip = 0;
}
inline int clipSegmentToLine(ClipVertex vOut[2], const ClipVertex vIn[2], const Vec2& normal, float offset) {
    int numOut = 0;
    float distance0 = dot(normal, vIn[0].v) - offset;
    float distance1 = dot(normal, vIn[1].v) - offset;
    if (distance0 <= 0.0f) vOut[numOut++] = vIn[0];
    if (distance1 <= 0.0f) vOut[numOut++] = vIn[1];
    if (distance0 * distance1 < 0.0f) {
        float interp = distance0 / (distance0 - distance1);
        vOut[numOut].v = vIn[0].v + interp * (vIn[1].v - vIn[0].v);
        if (distance0 > 0.0f)
            vOut[numOut].id = vIn[0].id;
        else
            vOut[numOut].id = vIn[1].id;
        ++numOut;
    }
    return numOut;
}
inline void collidePolygons(Manifold* m, const Shape* polyA, const Transform& xfA, const Shape* polyB, const Transform& xfB) {
    m->pointCount = 0;
    float totalRadius = polyA->radius + polyB->radius;
    int edgeA = 0;
    float separationA = findMaxSeparation(&edgeA, polyA, xfA, polyB, xfB);
    if (separationA > totalRadius) return;
    int edgeB = 0;
    float separationB = findMaxSeparation(&edgeB, polyB, xfB, polyA, xfA);
    if (separationB > totalRadius) return;
    const Shape *poly1, *poly2;
    Transform xf1, xf2;
    int edge1;
    uint8_t flip;
    const float k_relativeTol = 0.98f, k_absoluteTol = 0.001f;
    if (separationB > k_relativeTol * separationA + k_absoluteTol) {
        poly1 = polyB;
        poly2 = polyA;
        xf1 = xfB;
        xf2 = xfA;
        edge1 = edgeB;
        m->type = M_FACE_B;
        flip = 1;
    } else {
        poly1 = polyA;
        poly2 = polyB;
        xf1 = xfA;
        xf2 = xfB;
        edge1 = edgeA;
        m->type = M_FACE_A;
        flip = 0;
    }
    ClipVertex incidentEdge[2];
    findIncidentEdge(incidentEdge, poly1, xf1, edge1, poly2, xf2);
    int count1 = poly1->count;
    Vec2 v11 = poly1->v[edge1];
    Vec2 v12 = edge1 + 1 < count1 ? poly1->v[edge1 + 1] : poly1->v[0];
    Vec2 localTangent = v12 - v11;
    localTangent.normalize();
    Vec2 localNormal = cross(localTangent, 1.0f);
    Vec2 planePoint = 0.5f * (v11 + v12);
    Vec2 tangent = mul(xf1.R, localTangent);
    Vec2 normal = cross(tangent, 1.0f);
    v11 = mul(xf1, v11);
    v12 = mul(xf1, v12);
    float frontOffset = dot(normal, v11);
    float sideOffset1 = -dot(tangent, v11) + totalRadius;
    float sideOffset2 = dot(tangent, v12) + totalRadius;
    ClipVertex clipPoints1[2], clipPoints2[2];
    int np = clipSegmentToLine(clipPoints1, incidentEdge, -tangent, sideOffset1);
    if (np < 2) return;
    np = clipSegmentToLine(clipPoints2, clipPoints1, tangent, sideOffset2);
    if (np < 2) return;
    m->localNormal = localNormal;
    m->localPoint = planePoint;
    int pointCount = 0;
    for (int i = 0; i < kMaxManifoldPoints; ++i) {
        float separation = dot(normal, clipPoints2[i].v) - frontOffset;
        if (separation <= totalRadius) {
            ManifoldPoint* cp = m->points + pointCount;
            cp->localPoint = mulT(xf2, clipPoints2[i].v);
            cp->id = clipPoints2[i].id;
            cp->id.flip = flip;
            ++pointCount;
        }
    }
    m->pointCount = pointCount;
}

// ---- GJK distance and time of impact (A.7, A.9) ----
struct DistanceProxy {
    const Vec2* vertices = nullptr;
    int count = 0;
    float radius = 0;
    void set(const Shape* s) {
        if (s->type == CIRCLE) {
            vertices = &s->p;
            count = 1;
            radius = s->radius;
        } else {
            vertices = s->v;
            count = s->count;
            radius = s->radius;
        }
    }
    int getSupport(const Vec2& d) const {
        int bestIndex = 0;
        float bestValue = dot(vertices[0], d);
        for (int i = 1; i < count; ++i) {
            float value = dot(vertices[i], d);
            if (value > bestValue) {
                bestIndex = i;
                bestValue = value;
            }
        }
        return bestIndex;
    }
    const Vec2& getVertex(int i) const { return vertices[i]; }
};
struct SimplexCache {
    float metric = 0;
    int count = 0;
    int indexA[3] = {0, 0, 0}, indexB[3] = {0, 0, 0};
};
struct SimplexVertex {
    Vec2 wA, wB, w;
    float a = 0;
    int indexA = 0, indexB = 0;
};
struct Simplex {
    SimplexVertex v[3];
    int count = 0;
    float getMetric() const {
        switch (count) {
            case 1: return 0.0f;
            case 2: return distance(v[0].w, v[1].w);
            case 3: return cross(v[1].w - v[0].w, v[2].w - v[0].w);
            default: return 0.0f;
        }
    }
    void readCache(const SimplexCache* cache, const DistanceProxy* pA, const Transform& xfA, const DistanceProxy* pB, const Transform& xfB) {
        count = cache->count;
        for (int i = 0; i < count; ++i) {
            SimplexVertex* s = v + i;
            s->indexA = cache->indexA[i];
            s->indexB = cache->indexB[i];
            s->wA = mul(xfA, pA->getVertex(s->indexA));
            s->wB = mul(xfB, pB->getVertex(s->indexB));
            s->w = s->wB - s->wA;
            s->a = 0.0f;
        }
        if (count > 1) {
            float metric1 = cache->metric, metric2 = getMetric();
            if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < kEpsilon) count = 0;
        }
        if (count == 0) {
            SimplexVertex* s = v + 0;
            s->indexA = 0;
            s->indexB = 0;
            s->wA = mul(xfA, pA->getVertex(0));
            s->wB = mul(xfB, pB->getVertex(0));
            s->w = s->wB - s->wA;
            count = 1;
        }
    }
    void writeCache(SimplexCache* cache) const {
        cache->metric = getMetric();
        cache->count = count;
        for (int i = 0; i < count; ++i) {
            cache->indexA[i] = v[i].indexA;
            cache->indexB[i] = v[i].indexB;
        }
    }
    Vec2 getSearchDirection() const {
        if (count == 1) return -v[0].w;
        Vec2 e12 = v[1].w - v[0].w;
        float sgn = cross(e12, -v[0].w);
        if (sgn > 0.0f) return cross(1.0f, e12);
        return cross(e12, 1.0f);
    }
    Vec2 getClosestPoint() const {
        switch (count) {
            case 1: return v[0].w;
            case 2: return v[0].a * v[0].w + v[1].a * v[1].w;
            default: return Vec2();
        }
    }
    void getWitnessPoints(Vec2* pA, Vec2* pB) const {
        switch (count) {
            case 1:
                *pA = v[0].wA;
                *pB = v[0].wB;
                break;
            case 2:
                *pA = v[0].a * v[0].wA + v[1].a * v[1].wA;
                *pB = v[0].a * v[0].wB + v[1].a * v[1].wB;
                break;
            case 3:
                *pA = v[0].a * v[0].wA + v[1].a * v[1].wA + v[2].a * v[2].wA;
                *pB = *pA;
                break;
        }
    }
    void solve2() {
        Vec2 w1 = v[0].w, w2 = v[1].w;
        Vec2 e12 = w2 - w1;
        float d12_2 = -dot(w1, e12);
        if (d12_2 <= 0.0f) {
            v[0].a = 1.0f;
            count = 1;
            return;
        }
        float d12_1 = dot(w2, e12);
        if (d12_1 <= 0.0f) {
            v[1].a = 1.0f;
            count = 1;
            v[0] = v[1];
            return;
        }
        float inv = 1.0f / (d12_1 + d12_2);
        v[0].a = d12_1 * inv;
        v[1].a = d12_2 * inv;
        count = 2;
    }
    void solve3() {
        Vec2 w1 = v[0].w, w2 = v[1].w, w3 = v[2].w;
        Vec2 e12 = w2 - w1;
        float w1e12 = dot(w1, e12), w2e12 = dot(w2, e12);
        float d12_1 = w2e12, d12_2 = -w1e12;
        Vec2 e13 = w3 - w1;
        float w1e13 = dot(w1, e13), w3e13 = dot(w3, e13);
        float d13_1 = w3e13, d13_2 = -w1e13;
        Vec2 e23 = w3 - w2;
        float w2e23 = dot(w2, e23), w3e23 = dot(w3, e23);
        float d23_1 = w3e23, d23_2 = -w2e23;
        float n123 = cross(e12, e13);
        float d123_1 = n123 * cross(w2, w3);
        float d123_2 = n123 * cross(w3, w1);
        float d123_3 = n123 * cross(w1, w2);
        if (d12_2 <= 0.0f && d13_2 <= 0.0f) {
            v[0].a = 1.0f;
            count = 1;
            return;
        }
        if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {
            float inv = 1.0f / (d12_1 + d12_2);
            v[0].a = d12_1 * inv;
            v[1].a = d12_2 * inv;
            count = 2;
            return;
        }
        if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {
            float inv = 1.0f / (d13_1 + d13_2);
            v[0].a = d13_1 * inv;
            v[2].a = d13_2 * inv;
            count = 2;
            v[1] = v[2];
            return;
        }
        if (d12_1 <= 0.0f && d23_2 <= 0.0f) {
            v[1].a = 1.0f;
            count = 1;
            v[0] = v[1];
            return;
        }
        if (d13_1 <= 0.0f && d23_1 <= 0.0f) {
            v[2].a = 1.0f;
            count = 1;
            v[0] = v[2];
            return;
        }
        if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {
            float inv = 1.0f / (d23_1 + d23_2);
            v[1].a = d23_1 * inv;
            v[2].a = d23_2 * inv;
            count = 2;
            v[0] = v[2];
            return;
        }
        float inv = 1.0f / (d123_1 + d123_2 + d123_3);
        v[0].a = d123_1 * inv;
        v[1].a = d123_2 * inv;
        v[2].a = d123_3 * inv;
        count = 3;
    }
};
// b2Distance with useRadii = false (the only use on this path is from timeOfImpact)
inline float gjkDistance(SimplexCache* cache, const DistanceProxy* proxyA, const Transform& xfA, const DistanceProxy* proxyB, const Transform& xfB) {
    Simplex simplex;
    simplex.readCache(cache, proxyA, xfA, proxyB, xfB);
    SimplexVertex* vertices = simplex.v;
    const int k_maxIters = 20;
    int saveA[3], saveB[3], saveCount = 0;
    int iter = 0;
    while (iter < k_maxIters) {
        saveCount = simplex.count;
        for (int i = 0; i < saveCount; ++i) {
            saveA[i] = vertices[i].indexA;
            saveB[i] = vertices[i].indexB;
        }
        switch (simplex.count) {
            case 1: break;
            case 2: simplex.solve2(); break;
            case 3: simplex.solve3(); break;
        }
        if (simplex.count == 3) break;
        Vec2 d = simplex.getSearchDirection();
        if (d.lengthSquared() < kEpsilon * kEpsilon) break;
        SimplexVertex* vertex = vertices + simplex.count;
        vertex->indexA = proxyA->getSupport(mulT(xfA.R, -d));
        vertex->wA = mul(xfA, proxyA->getVertex(vertex->indexA));
        vertex->indexB = proxyB->getSupport(mulT(xfB.R, d));
        vertex->wB = mul(xfB, proxyB->getVertex(vertex->indexB));
        vertex->w = vertex->wB - vertex->wA;
        ++iter;
        bool duplicate = false;
        for (int i = 0; i < saveCount; ++i) {
            if (vertex->indexA == saveA[i] && vertex->indexB == saveB[i]) {
                duplicate = true;
                break;
            }
        }
        if (duplicate) break;
        ++simplex.count;
    }
    Vec2 pA, pB;
    simplex.getWitnessPoints(&pA, &pB);
    float dist = distance(pA, pB);
    simplex.writeCache(cache);
    return dist;
}

enum TOIState { TOI_UNKNOWN = 0, TOI_FAILED, TOI_OVERLAPPED, TOI_TOUCHING, TOI_SEPARATED };
struct SeparationFunction {
    enum Type { POINTS, FACE_A, FACE_B };
    const DistanceProxy *proxyA, *proxyB;
    Sweep sweepA, sweepB;
    Type type;
    Vec2 localPoint, axis;
    const Trig* tr;
    float initialize(const SimplexCache* cache, const DistanceProxy* pA, const Sweep& sA, const DistanceProxy* pB, const Sweep& sB, float t1, const Trig* trig) {
        proxyA = pA;
        proxyB = pB;
        tr = trig;
        int count = cache->count;
        sweepA = sA;
        sweepB = sB;
        Transform xfA, xfB;
        sweepA.getTransform(&xfA, t1, *tr);
        sweepB.getTransform(&xfB, t1, *tr);
        if (count == 1) {
            type = POINTS;
            Vec2 localPointA = proxyA->getVertex(cache->indexA[0]);
            Vec2 localPointB = proxyB->getVertex(cache->indexB[0]);
            Vec2 pointA = mul(xfA, localPointA), pointB = mul(xfB, localPointB);
            axis = pointB - pointA;
            return axis.normalize();
        } else if (cache->indexA[0] == cache->indexA[1]) {
            type = FACE_B;
            Vec2 localPointB1 = proxyB->getVertex(cache->indexB[0]);
            Vec2 localPointB2 = proxyB->getVertex(cache->indexB[1]);
            axis = cross(localPointB2 - localPointB1, 1.0f);
            axis.normalize();
            Vec2 normal = mul(xfB.R, axis);
            localPoint = 0.5f * (localPointB1 + localPointB2);
            Vec2 pointB = mul(xfB, localPoint);
            Vec2 localPointA = proxyA->getVertex(cache->indexA[0]);
            Vec2 pointA = mul(xfA, localPointA);
            float s = dot(pointA - pointB, normal);
            if (s < 0.0f) {
                axis = -axis;
                s = -s;
            }
            return s;
        } else {
            type = FACE_A;
            Vec2 localPointA1 = proxyA->getVertex(cache->indexA[0]);
            Vec2 localPointA2 = proxyA->getVertex(cache->indexA[1]);
            axis = cross(localPointA2 - localPointA1, 1.0f);
            axis.normalize();
            Vec2 normal = mul(xfA.R, axis);
            localPoint = 0.5f * (localPointA1 + localPointA2);
            Vec2 pointA = mul(xfA, localPoint);
            Vec2 localPointB = proxyB->getVertex(cache->indexB[0]);
            Vec2 pointB = mul(xfB, localPointB);
            float s = dot(pointB - pointA, normal);
            if (s < 0.0f) {
                axis = -axis;
                s = -s;
            }
            return s;
        }
    }
    float findMinSeparation(int* indexA, int* indexB, float t) const {
        Transform xfA, xfB;
        sweepA.getTransform(&xfA, t, *tr);
        sweepB.getTransform(&xfB, t, *tr);
        switch (type) {
            case POINTS: {
                Vec2 axisA = mulT(xfA.R, axis), axisB = mulT(xfB.R, -axis);
                *indexA = proxyA->getSupport(axisA);
                *indexB = proxyB->getSupport(axisB);
                Vec2 pointA = mul(xfA, proxyA->getVertex(*indexA));
                Vec2 pointB = mul(xfB, proxyB->getVertex(*indexB));
                return dot(pointB - pointA, axis);
            }
            case FACE_A: {
                Vec2 normal = mul(xfA.R, axis);
                Vec2 pointA = mul(xfA, localPoint);
                Vec2 axisB = mulT(xfB.R, -normal);
                *indexA = -1;
                *indexB = proxyB->getSupport(axisB);
                Vec2 pointB = mul(xfB, proxyB->getVertex(*indexB));
                return dot(pointB - pointA, normal);
            }
            default: {
                Vec2 normal = mul(xfB.R, axis);
                Vec2 pointB = mul(xfB, localPoint);
                Vec2 axisA = mulT(xfA.R, -normal);
                *indexB = -1;
                *indexA = proxyA->getSupport(axisA);
                Vec2 pointA = mul(xfA, proxyA->getVertex(*indexA));
                return dot(pointA - pointB, normal);
            }
        }
    }
    float evaluate(int indexA, int indexB, float t) const {
        Transform xfA, xfB;
        sweepA.getTransform(&xfA, t, *tr);
        sweepB.getTransform(&xfB, t, *tr);
        switch (type) {
            case POINTS: {
                Vec2 pointA = mul(xfA, proxyA->getVertex(indexA));
                Vec2 pointB = mul(xfB, proxyB->getVertex(indexB));
                return dot(pointB - pointA, axis);
            }
            case FACE_A: {
                Vec2 normal = mul(xfA.R, axis);
                Vec2 pointA = mul(xfA, localPoint);
                Vec2 pointB = mul(xfB, proxyB->getVertex(indexB));
                return dot(pointB - pointA, normal);
            }
            default: {
                Vec2 normal = mul(xfB.R, axis);
                Vec2 pointB = mul(xfB, localPoint);
                Vec2 pointA = mul(xfA, proxyA->getVertex(indexA));
                return dot(pointA - pointB, normal);
            }
        }
    }
};
inline void timeOfImpact(int* outState, float* outT, const DistanceProxy* proxyA, Sweep sweepA, const DistanceProxy* proxyB, Sweep sweepB, float tMax, const Trig& tr) {
    *outState = TOI_UNKNOWN;
    *outT = tMax;
    sweepA.normalize();
    sweepB.normalize();
    float totalRadius = proxyA->radius + proxyB->radius;
    float target = maxf(kLinearSlop, totalRadius - 3.0f * kLinearSlop);
    float tolerance = 0.25f * kLinearSlop;
    float t1 = 0.0f;
    const int k_maxIterations = 20;
    int iter = 0;
    SimplexCache cache;
    cache.count = 0;
    for (;;) {
        Transform xfA, xfB;
        sweepA.getTransform(&xfA, t1, tr);
        sweepB.getTransform(&xfB, t1, tr);
        float dist = gjkDistance(&cache, proxyA, xfA, proxyB, xfB);
        if (dist <= 0.0f) {
            *outState = TOI_OVERLAPPED;
            *outT = 0.0f;
            break;
        }
        if (dist < target + tolerance) {
            *outState = TOI_TOUCHING;
            *outT = t1;
            break;
        }
        SeparationFunction fcn;
        fcn.initialize(&cache, proxyA, sweepA, proxyB, sweepB, t1, &tr);
        bool done = false;
        float t2 = tMax;
        int pushBackIter = 0;
        for (;;) {
            int indexA, indexB;
            float s2 = fcn.findMinSeparation(&indexA, &indexB, t2);
            if (s2 > target + tolerance) {
                *outState = TOI_SEPARATED;
                *outT = tMax;
                done = true;
                break;
            }
            if (s2 > target - tolerance) {
                t1 = t2;
                break;
            }
            float s1 = fcn.evaluate(indexA, indexB, t1);
            if (s1 < target - tolerance) {
                *outState = TOI_FAILED;
                *outT = t1;
                done = true;
                break;
            }
            if (s1 <= target + tolerance) {
                *outState = TOI_TOUCHING;
                *outT = t1;
                done = true;
                break;
            }
            int rootIterCount = 0;
            float a1 = t1, a2 = t2;
            for (;;) {
                float t;
                if (rootIterCount & 1)
                    t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
                else
                    t = 0.5f * (a1 + a2);
                float s = fcn.evaluate(indexA, indexB, t);
                if (std::fabs(s - target) < tolerance) {
                    t2 = t;
                    break;
                }
                if (s > target) {
                    a1 = t;
                    s1 = s;
                } else {
                    a2 = t;
                    s2 = s;
                }
                ++rootIterCount;
                if (rootIterCount == 50) break;
            }
            ++pushBackIter;
            if (pushBackIter == kMaxPolygonVertices) break;
        }
        ++iter;
        if (done) break;
        if (iter == k_maxIterations) {
            *outState = TOI_FAILED;
            *outT = t1;
            break;
        }
    }
}

// ---- dynamics ----
enum BodyType { STATIC_BODY = 0, DYNAMIC_BODY = 2 };
enum UserKind { U_WALL = 0, U_HIDDEN = 1, U_ROBOT = 2, U_PUCK = 3 };  // Fixture userData class (PointSensor.java:181-194)

struct Body;
struct Contact;
struct Fixture {
    Body* body = nullptr;
    Shape shape;
    float density = 0, friction = 0.2f, restitution = 0;
    uint16_t categoryBits = 0x0001, maskBits = 0xFFFF;
    int16_t groupIndex = 0;
    int proxyId = -1;   // fixture creation order == broad-phase proxy order (A.3)
    AABB aabb;          // Fixture.m_aabb: swept tight AABB, read by PointSensor.possibleOverlap
    AABB fat;           // DynamicTree node AABB
    int userKind = U_WALL;
    int colour = 0;     // puck colour index when userKind == U_PUCK
    Fixture* next = nullptr;
};
struct ContactEdge {
    Body* other = nullptr;
    Contact* contact = nullptr;
    ContactEdge *prev = nullptr, *next = nullptr;
};
struct Contact {
    Fixture *fixtureA = nullptr, *fixtureB = nullptr;
    Manifold manifold;
    bool touching = false, islandFlag = false;
    int toiCount = 0;
    Contact *prev = nullptr, *next = nullptr;
    ContactEdge nodeA, nodeB;
    void evaluate(const Transform& xfA, const Transform& xfB) {
        const Shape *a = &fixtureA->shape, *b = &fixtureB->shape;
        if (a->type == CIRCLE && b->type == CIRCLE)
            collideCircles(&manifold, a, xfA, b, xfB);
        else if (a->type == POLYGON && b->type == CIRCLE)
            collidePolygonAndCircle(&manifold, a, xfA, b, xfB);
        else
            collidePolygons(&manifold, a, xfA, b, xfB);
    }
    void update();
};
struct Body {
    int type = STATIC_BODY;
    Transform xf;
    Sweep sweep;
    Vec2 linearVelocity;
    float angularVelocity = 0;
    Vec2 force;
    float torque = 0;
    float mass = 0, invMass = 0, I = 0, invI = 0;
    float linearDamping = 0, angularDamping = 0;
    Fixture* fixtureList = nullptr;
    ContactEdge* contactList = nullptr;
    Body *prev = nullptr, *next = nullptr;
    bool islandFlag = false, toiFlag = false;
    int index = 0;      // creation order
    int userKind = U_WALL;
    int colour = 0;
    const Trig* tr = nullptr;

    void synchronizeTransform() {
        xf.R.set(sweep.a, *tr);
        xf.position = sweep.c - mul(xf.R, sweep.localCenter);
    }
    Vec2 getWorldPoint(const Vec2& p) const { return mul(xf, p); }
    Vec2 getWorldVector(const Vec2& v) const { return mul(xf.R, v); }
    void applyForce(const Vec2& f, const Vec2& point) {
        if (type != DYNAMIC_BODY) return;
        force += f;
        torque += cross(point - sweep.c, f);
    }
    void applyTorque(float t) {
        if (type != DYNAMIC_BODY) return;
        torque += t;
    }
    void advance(float t) {
        sweep.advance(t);
        sweep.c = sweep.c0;
        sweep.a = sweep.a0;
        synchronizeTransform();
    }
    void resetMassData() {
        mass = 0;
        invMass = 0;
        I = 0;
        invI = 0;
        sweep.localCenter = Vec2();
        if (type == STATIC_BODY) {
            sweep.c0 = sweep.c = xf.position;
            return;
        }
        Vec2 center;
        for (Fixture* f = fixtureList; f; f = f->next) {
            if (f->density == 0.0f) continue;
            MassData md;
            f->shape.computeMass(&md, f->density);
            mass += md.mass;
            center += md.mass * md.center;
            I += md.I;
        }
        if (mass > 0.0f) {
            invMass = 1.0f / mass;
            center *= invMass;
        } else {
            mass = 1.0f;
            invMass = 1.0f;
        }
        if (I > 0.0f) {
            I -= mass * dot(center, center);
            invI = 1.0f / I;
        } else {
            I = 0;
            invI = 0;
        }
        Vec2 oldCenter = sweep.c;
        sweep.localCenter = center;
        sweep.c0 = sweep.c = mul(xf, sweep.localCenter);
        linearVelocity += cross(angularVelocity, sweep.c - oldCenter);
    }
};
inline void Contact::update() {
    Manifold oldManifold = manifold;
    evaluate(fixtureA->body->xf, fixtureB->body->xf);
    touching = manifold.pointCount > 0;
    for (int i = 0; i < manifold.pointCount; ++i) {
        ManifoldPoint* mp2 = manifold.points + i;
        mp2->normalImpulse = 0.0f;
        mp2->tangentImpulse = 0.0f;
        uint32_t id2 = mp2->id.key();
        for (int j = 0; j < oldManifold.pointCount; ++j) {
            ManifoldPoint* mp1 = oldManifold.points + j;
            if (mp1->id.key() == id2) {
                mp2->normalImpulse = mp1->normalImpulse;
                mp2->tangentImpulse = mp1->tangentImpulse;
                break;
            }
        }
    }
}

struct BodyDef {
    int type = STATIC_BODY;
    Vec2 position;
    float angle = 0;
    float linearDamping = 0, angularDamping = 0;
};
struct FixtureDef {
    Shape shape;
    float density = 0, friction = 0.2f, restitution = 0;
    uint16_t categoryBits = 0x0001, maskBits = 0xFFFF;
    int userKind = U_WALL;
    int colour = 0;
};

struct ContactConstraintPoint {
    Vec2 localPoint, rA, rB;
    float normalImpulse, tangentImpulse, normalMass, tangentMass, velocityBias;
};
struct ContactConstraint {
    ContactConstraintPoint points[kMaxManifoldPoints];
    Vec2 localNormal, localPoint, normal;
    Mat22 normalMass, K;
    Body *bodyA, *bodyB;
    int type;
    float radius, friction, restitution;
    int pointCount;
    Manifold* manifold;
};

struct World {
    Switches sw;
    Trig tr;
    Body* bodyList = nullptr;
    Contact* contactList = nullptr;
    int bodyCount = 0, contactCount = 0, proxyCount = 0;
    std::vector<Fixture*> proxies;   // by proxyId
    std::vector<int> moveBuffer;
    bool newFixture = false;
    float inv_dt0 = 0;
    Vec2 gravity;
    std::vector<Body*> allBodies;    // creation order (ownership)
    // statistics of the last step (reported by the oracle API)
    int lastPositionIterations = 0, lastToiEvents = 0, lastIslands = 0;
    std::vector<uint32_t> lastOrder;   // (proxyA << 16 | proxyB) of every solved contact, in Gauss-Seidel order, island after island

    World() { tr.lut = sw.sincosLut; }
    void setSwitches(const Switches& s) {
        sw = s;
        tr.lut = s.sincosLut;
    }
    ~World() {
        for (Contact* c = contactList; c;) {
            Contact* n = c->next;
            delete c;
            c = n;
        }
        for (Body* b : allBodies) {
            for (Fixture* f = b->fixtureList; f;) {
                Fixture* n = f->next;
                delete f;
                f = n;
            }
            delete b;
        }
    }
    World(const World&) = delete;
    World& operator=(const World&) = delete;

    Body* createBody(const BodyDef& bd) {
        Body* b = new Body();
        b->tr = &tr;
        b->type = bd.type;
        b->xf.position = bd.position;
        b->xf.R.set(bd.angle, tr);
        b->sweep.localCenter = Vec2();
        b->sweep.a0 = b->sweep.a = bd.angle;
        b->sweep.c0 = b->sweep.c = mul(b->xf, b->sweep.localCenter);
        b->linearDamping = bd.linearDamping;
        b->angularDamping = bd.angularDamping;
        if (bd.type == DYNAMIC_BODY) {
            b->mass = 1.0f;
            b->invMass = 1.0f;
        }
        b->index = bodyCount;
        b->prev = nullptr;
        b->next = bodyList;
        if (bodyList) bodyList->prev = b;
        bodyList = b;
        ++bodyCount;
        allBodies.push_back(b);
        return b;
    }
    Fixture* createFixture(Body* b, const FixtureDef& def) {
        Fixture* f = new Fixture();
        f->body = b;
        f->shape = def.shape;
        f->density = def.density;
        f->friction = def.friction;
        f->restitution = def.restitution;
        f->categoryBits = def.categoryBits;
        f->maskBits = def.maskBits;
        f->userKind = def.userKind;
        f->colour = def.colour;
        // Fixture.createProxy: tight AABB at the current transform, fattened by aabbExtension; buffered as moved
        f->shape.computeAABB(&f->aabb, b->xf);
        f->fat = f->aabb;
        Vec2 r(kAabbExtension, kAabbExtension);
        f->fat.lower = f->aabb.lower - r;
        f->fat.upper = f->aabb.upper + r;
        f->proxyId = proxyCount++;
        proxies.push_back(f);
        moveBuffer.push_back(f->proxyId);
        f->next = b->fixtureList;
        b->fixtureList = f;
        if (f->density > 0.0f) b->resetMassData();
        newFixture = true;
        return f;
    }

    // ---- ContactManager ----
    static bool shouldCollide(const Fixture* a, const Fixture* b) {
        if (a->groupIndex == b->groupIndex && a->groupIndex != 0) return a->groupIndex > 0;
        return (a->maskBits & b->categoryBits) != 0 && (a->categoryBits & b->maskBits) != 0;
    }
    void addPair(Fixture* fixtureA, Fixture* fixtureB) {
        Body *bodyA = fixtureA->body, *bodyB = fixtureB->body;
        if (bodyA == bodyB) return;
        for (ContactEdge* edge = bodyB->contactList; edge; edge = edge->next) {
            if (edge->other == bodyA) {
                Fixture *fA = edge->contact->fixtureA, *fB = edge->contact->fixtureB;
                if (fA == fixtureA && fB == fixtureB) return;
                if (fA == fixtureB && fB == fixtureA) return;
            }
        }
        if (bodyB->type != DYNAMIC_BODY && bodyA->type != DYNAMIC_BODY) return;
        if (!shouldCollide(fixtureA, fixtureB)) return;
        // Contact.create: the type registry puts the polygon first in polygon-circle pairs
        if (fixtureA->shape.type == CIRCLE && fixtureB->shape.type == POLYGON) std::swap(fixtureA, fixtureB);
        insertContact(fixtureA, fixtureB);
    }
    Contact* insertContact(Fixture* fixtureA, Fixture* fixtureB) {
        Contact* c = new Contact();
        c->fixtureA = fixtureA;
        c->fixtureB = fixtureB;
        Body *bodyA = fixtureA->body, *bodyB = fixtureB->body;
        c->prev = nullptr;
        c->next = contactList;
        if (contactList) contactList->prev = c;
        contactList = c;
        c->nodeA.contact = c;
        c->nodeA.other = bodyB;
        c->nodeA.prev = nullptr;
        c->nodeA.next = bodyA->contactList;
        if (bodyA->contactList) bodyA->contactList->prev = &c->nodeA;
        bodyA->contactList = &c->nodeA;
        c->nodeB.contact = c;
        c->nodeB.other = bodyA;
        c->nodeB.prev = nullptr;
        c->nodeB.next = bodyB->contactList;
        if (bodyB->contactList) bodyB->contactList->prev = &c->nodeB;
        bodyB->contactList = &c->nodeB;
        ++contactCount;
        return c;
    }
    void destroyContact(Contact* c) {
        Body *bodyA = c->fixtureA->body, *bodyB = c->fixtureB->body;
        if (c->prev) c->prev->next = c->next;
        if (c->next) c->next->prev = c->prev;
        if (c == contactList) contactList = c->next;
        if (c->nodeA.prev) c->nodeA.prev->next = c->nodeA.next;
        if (c->nodeA.next) c->nodeA.next->prev = c->nodeA.prev;
        if (&c->nodeA == bodyA->contactList) bodyA->contactList = c->nodeA.next;
        if (c->nodeB.prev) c->nodeB.prev->next = c->nodeB.next;
        if (c->nodeB.next) c->nodeB.next->prev = c->nodeB.prev;
        if (&c->nodeB == bodyB->contactList) bodyB->contactList = c->nodeB.next;
        delete c;
        --contactCount;
    }
    void clearContacts() {
        while (contactList) destroyContact(contactList);
    }
    // BroadPhase.updatePairs: every moved proxy queries all fat AABBs (the dynamic tree only accelerates
    // this query; the pair set, its sort order and its de-duplication are what A.3 specifies).
    void findNewContacts() {
        std::vector<std::pair<int, int>> pairs;
        for (int q : moveBuffer) {
            const AABB& fatQ = proxies[q]->fat;
            for (int p = 0; p < proxyCount; ++p) {
                if (p == q) continue;
                if (testOverlap(proxies[p]->fat, fatQ)) pairs.push_back(std::make_pair(std::min(p, q), std::max(p, q)));
            }
        }
        moveBuffer.clear();
        std::sort(pairs.begin(), pairs.end());
        size_t i = 0;
        while (i < pairs.size()) {
            std::pair<int, int> primary = pairs[i];
            addPair(proxies[primary.first], proxies[primary.second]);
            ++i;
            while (i < pairs.size() && pairs[i] == primary) ++i;
        }
    }
    void collide() {
        Contact* c = contactList;
        while (c) {
            if (!testOverlap(c->fixtureA->fat, c->fixtureB->fat)) {
                Contact* nuke = c;
                c = c->next;
                destroyContact(nuke);
                continue;
            }
            c->update();
            c = c->next;
        }
    }
    // Fixture.synchronize(broadPhase, xf1, xf2) + BroadPhase.moveProxy / DynamicTree.moveProxy
    void synchronizeFixture(Fixture* f, const Transform& xf1, const Transform& xf2) {
        AABB aabb1, aabb2;
        f->shape.computeAABB(&aabb1, xf1);
        f->shape.computeAABB(&aabb2, xf2);
        f->aabb.combine(aabb1, aabb2);
        Vec2 displacement = xf2.position - xf1.position;
        // DynamicTree.moveProxy
        if (f->fat.contains(f->aabb)) return;
        AABB nb = f->aabb;
        Vec2 r(kAabbExtension, kAabbExtension);
        nb.lower = nb.lower - r;
        nb.upper = nb.upper + r;
        Vec2 d = kAabbMultiplier * displacement;
        if (d.x < 0.0f) nb.lower.x += d.x; else nb.upper.x += d.x;
        if (d.y < 0.0f) nb.lower.y += d.y; else nb.upper.y += d.y;
        f->fat = nb;
        moveBuffer.push_back(f->proxyId);
    }
    void synchronizeFixtures(Body* b) {
        Transform xf1;
        xf1.R.set(b->sweep.a0, tr);
        xf1.position = b->sweep.c0 - mul(xf1.R, b->sweep.localCenter);
        for (Fixture* f = b->fixtureList; f; f = f->next) synchronizeFixture(f, xf1, b->xf);
    }
    // Body.setTransform (JBox2D 2.1.2.2 Body.setTransform == Box2D v2.1.2 b2Body::SetTransform; called by
    // src/arena/Arena.java:346,360,369,371 outside World.step, so the world is not locked): new transform, sweep
    // collapsed onto it (c0 = c, a0 = a), every fixture re-synchronised with xf1 == xf2 (zero displacement, so a
    // proxy that left its fat AABB gets the tight box + aabbExtension), then ContactManager.findNewContacts.
    // Velocities and the cached contacts stay: contacts whose fat AABBs no longer overlap die in the next collide().
    void setTransform(Body* b, const Vec2& position, float angle) {
        b->xf.R.set(angle, tr);
        b->xf.position = position;
        b->sweep.c0 = b->sweep.c = mul(b->xf, b->sweep.localCenter);
        b->sweep.a0 = b->sweep.a = angle;
        for (Fixture* f = b->fixtureList; f; f = f->next) synchronizeFixture(f, b->xf, b->xf);
        findNewContacts();
    }

    // ---- ContactSolver + Island.solve (A.6) ----
    void solveIsland(std::vector<Body*>& bodies, std::vector<Contact*>& contacts, float dt, float dtRatio, int velocityIterations, int positionIterations) {
        for (Body* b : bodies) {
            if (b->type != DYNAMIC_BODY) continue;
            b->linearVelocity += dt * (gravity + b->invMass * b->force);
            b->angularVelocity += dt * b->invI * b->torque;
            b->linearVelocity *= clampf(1.0f - dt * b->linearDamping, 0.0f, 1.0f);
            b->angularVelocity *= clampf(1.0f - dt * b->angularDamping, 0.0f, 1.0f);
        }
        int contactCountI = (int)contacts.size();
        if (sw.partitionStaticLast) {
            int i1 = -1;
            for (int i2 = 0; i2 < contactCountI; ++i2) {
                Body *bodyA = contacts[i2]->fixtureA->body, *bodyB = contacts[i2]->fixtureB->body;
                bool nonStatic = bodyA->type != STATIC_BODY && bodyB->type != STATIC_BODY;
                if (nonStatic) {
                    ++i1;
                    std::swap(contacts[i1], contacts[i2]);
                }
            }
        }
        for (int i = 0; i < contactCountI; ++i)
            lastOrder.push_back(((uint32_t)contacts[i]->fixtureA->proxyId << 16) | (uint32_t)contacts[i]->fixtureB->proxyId);
        // ContactSolver constructor
        std::vector<ContactConstraint> cons(contactCountI);
        for (int i = 0; i < contactCountI; ++i) {
            Contact* contact = contacts[i];
            Fixture *fixtureA = contact->fixtureA, *fixtureB = contact->fixtureB;
            float radiusA = fixtureA->shape.radius, radiusB = fixtureB->shape.radius;
            Body *bodyA = fixtureA->body, *bodyB = fixtureB->body;
            Manifold* manifold = &contact->manifold;
            float friction = std::sqrt(fixtureA->friction * fixtureB->friction);
            float restitution = fixtureA->restitution > fixtureB->restitution ? fixtureA->restitution : fixtureB->restitution;
            Vec2 vA = bodyA->linearVelocity, vB = bodyB->linearVelocity;
            float wA = bodyA->angularVelocity, wB = bodyB->angularVelocity;
            WorldManifold worldManifold;
            worldManifold.initialize(manifold, bodyA->xf, radiusA, bodyB->xf, radiusB);
            ContactConstraint* cc = &cons[i];
            cc->bodyA = bodyA;
            cc->bodyB = bodyB;
            cc->manifold = manifold;
            cc->normal = worldManifold.normal;
            cc->pointCount = manifold->pointCount;
            cc->friction = friction;
            cc->restitution = restitution;
            cc->localNormal = manifold->localNormal;
            cc->localPoint = manifold->localPoint;
            cc->radius = radiusA + radiusB;
            cc->type = manifold->type;
            for (int j = 0; j < cc->pointCount; ++j) {
                ManifoldPoint* cp = manifold->points + j;
                ContactConstraintPoint* ccp = cc->points + j;
                ccp->normalImpulse = dtRatio * cp->normalImpulse;
                ccp->tangentImpulse = dtRatio * cp->tangentImpulse;
                ccp->localPoint = cp->localPoint;
                ccp->rA = worldManifold.points[j] - bodyA->sweep.c;
                ccp->rB = worldManifold.points[j] - bodyB->sweep.c;
                float rnA = cross(ccp->rA, cc->normal), rnB = cross(ccp->rB, cc->normal);
                rnA *= rnA;
                rnB *= rnB;
                float kNormal = bodyA->invMass + bodyB->invMass + bodyA->invI * rnA + bodyB->invI * rnB;
                ccp->normalMass = 1.0f / kNormal;
                Vec2 tangent = cross(cc->normal, 1.0f);
                float rtA = cross(ccp->rA, tangent), rtB = cross(ccp->rB, tangent);
                rtA *= rtA;
                rtB *= rtB;
                float kTangent = bodyA->invMass + bodyB->invMass + bodyA->invI * rtA + bodyB->invI * rtB;
                ccp->tangentMass = 1.0f / kTangent;
                ccp->velocityBias = 0.0f;
                float vRel = dot(cc->normal, vB + cross(wB, ccp->rB) - vA - cross(wA, ccp->rA));
                if (vRel < -kVelocityThreshold) ccp->velocityBias = -cc->restitution * vRel;
            }
            if (cc->pointCount == 2) {
                ContactConstraintPoint *ccp1 = cc->points + 0, *ccp2 = cc->points + 1;
                float invMassA = bodyA->invMass, invIA = bodyA->invI, invMassB = bodyB->invMass, invIB = bodyB->invI;
                float rn1A = cross(ccp1->rA, cc->normal), rn1B = cross(ccp1->rB, cc->normal);
                float rn2A = cross(ccp2->rA, cc->normal), rn2B = cross(ccp2->rB, cc->normal);
                float k11 = invMassA + invMassB + invIA * rn1A * rn1A + invIB * rn1B * rn1B;
                float k22 = invMassA + invMassB + invIA * rn2A * rn2A + invIB * rn2B * rn2B;
                float k12 = invMassA + invMassB + invIA * rn1A * rn2A + invIB * rn1B * rn2B;
                const float k_maxConditionNumber = 100.0f;
                if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12)) {
                    cc->K.col1.set(k11, k12);
                    cc->K.col2.set(k12, k22);
                    cc->normalMass = cc->K.getInverse();
                } else {
                    cc->pointCount = 1;
                }
            }
        }
        // warm start
        for (int i = 0; i < contactCountI; ++i) {
            ContactConstraint* c = &cons[i];
            Body *bodyA = c->bodyA, *bodyB = c->bodyB;
            float invMassA = bodyA->invMass, invIA = bodyA->invI, invMassB = bodyB->invMass, invIB = bodyB->invI;
            Vec2 normal = c->normal, tangent = cross(normal, 1.0f);
            for (int j = 0; j < c->pointCount; ++j) {
                ContactConstraintPoint* ccp = c->points + j;
                Vec2 P = ccp->normalImpulse * normal + ccp->tangentImpulse * tangent;
                bodyA->angularVelocity -= invIA * cross(ccp->rA, P);
                bodyA->linearVelocity -= invMassA * P;
                bodyB->angularVelocity += invIB * cross(ccp->rB, P);
                bodyB->linearVelocity += invMassB * P;
            }
        }
        for (int it = 0; it < velocityIterations; ++it) solveVelocityConstraints(cons);
        // store impulses
        for (int i = 0; i < contactCountI; ++i) {
            ContactConstraint* c = &cons[i];
            for (int j = 0; j < c->pointCount; ++j) {
                c->manifold->points[j].normalImpulse = c->points[j].normalImpulse;
                c->manifold->points[j].tangentImpulse = c->points[j].tangentImpulse;
            }
        }
        // integrate positions
        for (Body* b : bodies) {
            if (b->type == STATIC_BODY) continue;
            Vec2 translation = dt * b->linearVelocity;
            if (dot(translation, translation) > kMaxTranslationSquared) {
                float ratio = kMaxTranslation / translation.length();
                b->linearVelocity *= ratio;
            }
            float rotation = dt * b->angularVelocity;
            if (rotation * rotation > kMaxRotationSquared) {
                float ratio = kMaxRotation / std::fabs(rotation);
                b->angularVelocity *= ratio;
            }
            b->sweep.c0 = b->sweep.c;
            b->sweep.a0 = b->sweep.a;
            b->sweep.c += dt * b->linearVelocity;
            b->sweep.a += dt * b->angularVelocity;
            b->synchronizeTransform();
        }
        int iters = 0;
        for (int i = 0; i < positionIterations; ++i) {
            ++iters;
            bool contactsOkay = solvePositionConstraints(cons, kContactBaumgarte);
            if (contactsOkay) break;
        }
        lastPositionIterations = std::max(lastPositionIterations, iters);
    }
    void solveVelocityConstraints(std::vector<ContactConstraint>& cons) {
        for (size_t i = 0; i < cons.size(); ++i) {
            ContactConstraint* c = &cons[i];
            Body *bodyA = c->bodyA, *bodyB = c->bodyB;
            float wA = bodyA->angularVelocity, wB = bodyB->angularVelocity;
            Vec2 vA = bodyA->linearVelocity, vB = bodyB->linearVelocity;
            float invMassA = bodyA->invMass, invIA = bodyA->invI, invMassB = bodyB->invMass, invIB = bodyB->invI;
            Vec2 normal = c->normal, tangent = cross(normal, 1.0f);
            float friction = c->friction;
            for (int j = 0; j < c->pointCount; ++j) {
                ContactConstraintPoint* ccp = c->points + j;
                Vec2 dv = vB + cross(wB, ccp->rB) - vA - cross(wA, ccp->rA);
                float vt = dot(dv, tangent);
                float lambda = ccp->tangentMass * (-vt);
                float maxFriction = friction * ccp->normalImpulse;
                float newImpulse = clampf(ccp->tangentImpulse + lambda, -maxFriction, maxFriction);
                lambda = newImpulse - ccp->tangentImpulse;
                Vec2 P = lambda * tangent;
                vA -= invMassA * P;
                wA -= invIA * cross(ccp->rA, P);
                vB += invMassB * P;
                wB += invIB * cross(ccp->rB, P);
                ccp->tangentImpulse = newImpulse;
            }
            if (c->pointCount == 1) {
                ContactConstraintPoint* ccp = c->points + 0;
                Vec2 dv = vB + cross(wB, ccp->rB) - vA - cross(wA, ccp->rA);
                float vn = dot(dv, normal);
                float lambda = -ccp->normalMass * (vn - ccp->velocityBias);
                float newImpulse = maxf(ccp->normalImpulse + lambda, 0.0f);
                lambda = newImpulse - ccp->normalImpulse;
                Vec2 P = lambda * normal;
                vA -= invMassA * P;
                wA -= invIA * cross(ccp->rA, P);
                vB += invMassB * P;
                wB += invIB * cross(ccp->rB, P);
                ccp->normalImpulse = newImpulse;
            } else {
                ContactConstraintPoint *cp1 = c->points + 0, *cp2 = c->points + 1;
                Vec2 a(cp1->normalImpulse, cp2->normalImpulse);
                Vec2 dv1 = vB + cross(wB, cp1->rB) - vA - cross(wA, cp1->rA);
                Vec2 dv2 = vB + cross(wB, cp2->rB) - vA - cross(wA, cp2->rA);
                float vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
                Vec2 b;
                b.x = vn1 - cp1->velocityBias;
                b.y = vn2 - cp2->velocityBias;
                b -= mul(c->K, a);
                for (;;) {
                    Vec2 x = -mul(c->normalMass, b);
                    if (x.x >= 0.0f && x.y >= 0.0f) {
                        applyBlock(c, x, a, normal, vA, wA, vB, wB, invMassA, invIA, invMassB, invIB);
                        break;
                    }
                    x.x = -cp1->normalMass * b.x;
                    x.y = 0.0f;
                    vn1 = 0.0f;
                    vn2 = c->K.col1.y * x.x + b.y;
                    if (x.x >= 0.0f && vn2 >= 0.0f) {
                        applyBlock(c, x, a, normal, vA, wA, vB, wB, invMassA, invIA, invMassB, invIB);
                        break;
                    }
                    x.x = 0.0f;
                    x.y = -cp2->normalMass * b.y;
                    vn1 = c->K.col2.x * x.y + b.x;
                    vn2 = 0.0f;
                    if (x.y >= 0.0f && vn1 >= 0.0f) {
                        applyBlock(c, x, a, normal, vA, wA, vB, wB, invMassA, invIA, invMassB, invIB);
                        break;
                    }
                    x.x = 0.0f;
                    x.y = 0.0f;
                    vn1 = b.x;
                    vn2 = b.y;
                    if (vn1 >= 0.0f && vn2 >= 0.0f) {
                        applyBlock(c, x, a, normal, vA, wA, vB, wB, invMassA, invIA, invMassB, invIB);
                        break;
                    }
                    break;
                }
            }
            bodyA->linearVelocity = vA;
            bodyA->angularVelocity = wA;
            bodyB->linearVelocity = vB;
            bodyB->angularVelocity = wB;
        }
    }
    static void applyBlock(ContactConstraint* c, const Vec2& x, const Vec2& a, const Vec2& normal, Vec2& vA, float& wA, Vec2& vB, float& wB,
                           float invMassA, float invIA, float invMassB, float invIB) {
        ContactConstraintPoint *cp1 = c->points + 0, *cp2 = c->points + 1;
        Vec2 d = x - a;
        Vec2 P1 = d.x * normal, P2 = d.y * normal;
        vA -= invMassA * (P1 + P2);
        wA -= invIA * (cross(cp1->rA, P1) + cross(cp2->rA, P2));
        vB += invMassB * (P1 + P2);
        wB += invIB * (cross(cp1->rB, P1) + cross(cp2->rB, P2));
        cp1->normalImpulse = x.x;
        cp2->normalImpulse = x.y;
    }
    // PositionSolverManifold (also used by the TOI solver)
    static void positionManifold(int type, const Vec2& localNormal, const Vec2& localPoint, const Vec2& pointLocal, float radius, const Body* bodyA,
                                 const Body* bodyB, Vec2* normal, Vec2* point, float* separation) {
        switch (type) {
            case M_CIRCLES: {
                Vec2 pointA = bodyA->getWorldPoint(localPoint);
                Vec2 pointB = bodyB->getWorldPoint(pointLocal);
                if (distanceSquared(pointA, pointB) > kEpsilon * kEpsilon) {
                    *normal = pointB - pointA;
                    normal->normalize();
                } else {
                    normal->set(1.0f, 0.0f);
                }
                *point = 0.5f * (pointA + pointB);
                *separation = dot(pointB - pointA, *normal) - radius;
            } break;
            case M_FACE_A: {
                *normal = bodyA->getWorldVector(localNormal);
                Vec2 planePoint = bodyA->getWorldPoint(localPoint);
                Vec2 clipPoint = bodyB->getWorldPoint(pointLocal);
                *separation = dot(clipPoint - planePoint, *normal) - radius;
                *point = clipPoint;
            } break;
            case M_FACE_B: {
                *normal = bodyB->getWorldVector(localNormal);
                Vec2 planePoint = bodyB->getWorldPoint(localPoint);
                Vec2 clipPoint = bodyA->getWorldPoint(pointLocal);
                *separation = dot(clipPoint - planePoint, *normal) - radius;
                *point = clipPoint;
                *normal = -*normal;
            } break;
        }
    }
    bool solvePositionConstraints(std::vector<ContactConstraint>& cons, float baumgarte) {
        float minSeparation = 0.0f;
        for (size_t i = 0; i < cons.size(); ++i) {
            ContactConstraint* c = &cons[i];
            Body *bodyA = c->bodyA, *bodyB = c->bodyB;
            float invMassA, invIA, invMassB, invIB;
            if (sw.posSolveMassScaled) {
                invMassA = bodyA->mass * bodyA->invMass;
                invIA = bodyA->mass * bodyA->invI;
                invMassB = bodyB->mass * bodyB->invMass;
                invIB = bodyB->mass * bodyB->invI;
            } else {
                invMassA = bodyA->invMass;
                invIA = bodyA->invI;
                invMassB = bodyB->invMass;
                invIB = bodyB->invI;
            }
            for (int j = 0; j < c->pointCount; ++j) {
                Vec2 normal, point;
                float separation;
                positionManifold(c->type, c->localNormal, c->localPoint, c->points[j].localPoint, c->radius, bodyA, bodyB, &normal, &point, &separation);
                Vec2 rA = point - bodyA->sweep.c, rB = point - bodyB->sweep.c;
                minSeparation = minf(minSeparation, separation);
                float C = clampf(baumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
                float rnA = cross(rA, normal), rnB = cross(rB, normal);
                float K = invMassA + invMassB + invIA * rnA * rnA + invIB * rnB * rnB;
                float impulse = K > 0.0f ? -C / K : 0.0f;
                Vec2 P = impulse * normal;
                bodyA->sweep.c -= invMassA * P;
                bodyA->sweep.a -= invIA * cross(rA, P);
                bodyA->synchronizeTransform();
                bodyB->sweep.c += invMassB * P;
                bodyB->sweep.a += invIB * cross(rB, P);
                bodyB->synchronizeTransform();
            }
        }
        return minSeparation >= -1.5f * kLinearSlop;
    }

    // ---- World.solve: islands by DFS (A.5) ----
    void solve(float dt, float dtRatio, int velocityIterations, int positionIterations) {
        for (Body* b = bodyList; b; b = b->next) b->islandFlag = false;
        for (Contact* c = contactList; c; c = c->next) c->islandFlag = false;
        std::vector<Body*> stack(bodyCount);
        std::vector<Body*> islandBodies;
        std::vector<Contact*> islandContacts;
        lastPositionIterations = 0;
        lastIslands = 0;
        lastOrder.clear();
        for (Body* seed = bodyList; seed; seed = seed->next) {
            if (seed->islandFlag) continue;
            if (seed->type == STATIC_BODY) continue;
            islandBodies.clear();
            islandContacts.clear();
            int stackCount = 0;
            stack[stackCount++] = seed;
            seed->islandFlag = true;
            while (stackCount > 0) {
                Body* b = stack[--stackCount];
                islandBodies.push_back(b);
                if (b->type == STATIC_BODY) continue;
                for (ContactEdge* ce = b->contactList; ce; ce = ce->next) {
                    Contact* contact = ce->contact;
                    if (contact->islandFlag) continue;
                    if (!contact->touching) continue;
                    islandContacts.push_back(contact);
                    contact->islandFlag = true;
                    Body* other = ce->other;
                    if (other->islandFlag) continue;
                    stack[stackCount++] = other;
                    other->islandFlag = true;
                }
            }
            solveIsland(islandBodies, islandContacts, dt, dtRatio, velocityIterations, positionIterations);
            ++lastIslands;
            for (Body* b : islandBodies)
                if (b->type == STATIC_BODY) b->islandFlag = false;
        }
        for (Body* b = bodyList; b; b = b->next) {
            if (!b->islandFlag) continue;
            if (b->type == STATIC_BODY) continue;
            synchronizeFixtures(b);
        }
        findNewContacts();
    }

    // ---- World.solveTOI (A.7); no bullets on this path ----
    void solveTOIBody(Body* body) {
        Contact* toiContact = nullptr;
        float toi = 1.0f;
        bool found;
        int count, iter = 0;
        do {
            count = 0;
            found = false;
            for (ContactEdge* ce = body->contactList; ce; ce = ce->next) {
                if (ce->contact == toiContact) continue;
                Body* other = ce->other;
                if (other->type == DYNAMIC_BODY) continue;
                Contact* contact = ce->contact;
                if (contact->toiCount > 10) continue;
                Fixture *fixtureA = contact->fixtureA, *fixtureB = contact->fixtureB;
                Body *bodyA = fixtureA->body, *bodyB = fixtureB->body;
                DistanceProxy proxyA, proxyB;
                proxyA.set(&fixtureA->shape);
                proxyB.set(&fixtureB->shape);
                int state;
                float t;
                timeOfImpact(&state, &t, &proxyA, bodyA->sweep, &proxyB, bodyB->sweep, toi, tr);
                if (state == TOI_TOUCHING && t < toi) {
                    toiContact = contact;
                    toi = t;
                    found = true;
                }
                ++count;
            }
            ++iter;
        } while (found && count > 1 && iter < 50);
        if (toiContact == nullptr) {
            body->advance(1.0f);
            return;
        }
        ++lastToiEvents;
        body->advance(toi);
        toiContact->update();
        ++toiContact->toiCount;
        Contact* contacts[kMaxTOIContacts];
        count = 0;
        for (ContactEdge* ce = body->contactList; ce && count < kMaxTOIContacts; ce = ce->next) {
            Body* other = ce->other;
            if (other->type == DYNAMIC_BODY) continue;
            Contact* contact = ce->contact;
            if (contact != toiContact) contact->update();
            if (!contact->touching) continue;
            contacts[count] = contact;
            ++count;
        }
        // TOISolver: position-only correction of the TOI body (Baumgarte 0.75)
        const float k_toiBaumgarte = 0.75f;
        for (int i = 0; i < 20; ++i) {
            float minSeparation = 0.0f;
            for (int k = 0; k < count; ++k) {
                Contact* contact = contacts[k];
                Body *bodyA = contact->fixtureA->body, *bodyB = contact->fixtureB->body;
                const Manifold* m = &contact->manifold;
                float radius = contact->fixtureA->shape.radius + contact->fixtureB->shape.radius;
                float massA = bodyA->mass, massB = bodyB->mass;
                if (bodyA == body) massB = 0.0f; else massA = 0.0f;
                float invMassA = massA * bodyA->invMass, invIA = massA * bodyA->invI;
                float invMassB = massB * bodyB->invMass, invIB = massB * bodyB->invI;
                for (int j = 0; j < m->pointCount; ++j) {
                    Vec2 normal, point;
                    float separation;
                    positionManifold(m->type, m->localNormal, m->localPoint, m->points[j].localPoint, radius, bodyA, bodyB, &normal, &point, &separation);
                    Vec2 rA = point - bodyA->sweep.c, rB = point - bodyB->sweep.c;
                    minSeparation = minf(minSeparation, separation);
                    float C = clampf(k_toiBaumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
                    float rnA = cross(rA, normal), rnB = cross(rB, normal);
                    float K = invMassA + invMassB + invIA * rnA * rnA + invIB * rnB * rnB;
                    float impulse = K > 0.0f ? -C / K : 0.0f;
                    Vec2 P = impulse * normal;
                    bodyA->sweep.c -= invMassA * P;
                    bodyA->sweep.a -= invIA * cross(rA, P);
                    bodyA->synchronizeTransform();
                    bodyB->sweep.c += invMassB * P;
                    bodyB->sweep.a += invIB * cross(rB, P);
                    bodyB->synchronizeTransform();
                }
            }
            if (minSeparation >= -1.5f * kLinearSlop) break;
        }
    }
    void solveTOI() {
        lastToiEvents = 0;
        for (Contact* c = contactList; c; c = c->next) c->toiCount = 0;
        for (Body* b = bodyList; b; b = b->next) b->toiFlag = (!b->islandFlag || b->type == STATIC_BODY);
        for (Body* b = bodyList; b; b = b->next) {
            if (b->toiFlag) continue;
            solveTOIBody(b);
            b->toiFlag = true;
        }
    }

    // ---- World.step (A.8) ----
    void step(float dt, int velocityIterations, int positionIterations) {
        if (newFixture) {
            findNewContacts();
            newFixture = false;
        }
        float inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;
        float dtRatio = inv_dt0 * dt;
        collide();
        if (dt > 0.0f) solve(dt, dtRatio, velocityIterations, positionIterations);
        if (sw.toiEnabled && dt > 0.0f) solveTOI();
        if (dt > 0.0f) inv_dt0 = inv_dt;
        for (Body* b = bodyList; b; b = b->next) {  // clearForces (auto-clear is the default)
            b->force = Vec2();
            b->torque = 0.0f;
        }
    }
};

}  // namespace jb
