// puckswarm_b200 -- GJK distance and conservative-advancement time of impact with JBox2D 2.1.2 /
// Box2D v2.1.2 semantics (org.jbox2d.collision.Distance, TimeOfImpact).
// On this path it is only ever evaluated between one fixture of a dynamic body and one wall edge
// (World.solveTOI skips dynamic-dynamic pairs for non-bullets).
#pragma once
#include "ps_collide.cuh"

namespace ps {

struct Sweep {
    float lcx, lcy;          // localCenter
    float c0x, c0y, cx, cy;  // centre of mass at t = 0 and t = 1
    float a0, a;
};
PS_HD Xf sweepXf(const Trig& tr, const Sweep& s, float alpha) {
    Xf T;
    float px = (1.0f - alpha) * s.c0x + alpha * s.cx;
    float py = (1.0f - alpha) * s.c0y + alpha * s.cy;
    float angle = (1.0f - alpha) * s.a0 + alpha * s.a;
    T.c = psCos(tr, angle);
    T.s = psSin(tr, angle);
    T.px = px - (T.c * s.lcx + (-T.s) * s.lcy);
    T.py = py - (T.s * s.lcx + T.c * s.lcy);
    return T;
}
// How a body's transform depends on the sweep parameter: a static body never moves (the wall), a puck only translates
// (its rotation is never read: circle fixtures centred on the body origin, localCenter = 0), a robot needs the full
// Sweep.getTransform with MathUtils sin / cos. The shortcuts return exactly what the full formula would.
enum { SWEEP_STATIC = 0, SWEEP_TRANSLATE = 1, SWEEP_FULL = 2 };
PS_HD Xf sweepXfKind(const Trig& tr, const Sweep& s, int kind, const Xf& staticXf, float alpha) {
    if (kind == SWEEP_STATIC) return staticXf;
    if (kind == SWEEP_TRANSLATE) {
        Xf T;
        T.px = (1.0f - alpha) * s.c0x + alpha * s.cx;
        T.py = (1.0f - alpha) * s.c0y + alpha * s.cy;
        T.c = 1.0f;
        T.s = 0.0f;
        return T;
    }
    return sweepXf(tr, s, alpha);
}
PS_HD void sweepNormalize(Sweep& s) {
    float d = kTwoPi * floorf(s.a0 / kTwoPi);
    s.a0 -= d;
    s.a -= d;
}
PS_HD void sweepAdvance(Sweep& s, float t) {
    s.c0x = (1.0f - t) * s.c0x + t * s.cx;
    s.c0y = (1.0f - t) * s.c0y + t * s.cy;
    s.a0 = (1.0f - t) * s.a0 + t * s.a;
}

// DistanceProxy: a circle is one vertex (its centre), a polygon its vertices
struct DProxy {
    const float *vx, *vy;   // into the scene's shape table (no per-thread copy: a dynamically indexed local array would live in local memory)
    int count;
    float radius;
    PS_HD V2 v(int i) const { return mk(vx[i], vy[i]); }
    PS_HD int support(V2 d) const {
        int best = 0;
        float bestValue = dot(v(0), d);
        for (int i = 1; i < count; ++i) {
            float value = dot(v(i), d);
            if (value > bestValue) {
                best = i;
                bestValue = value;
            }
        }
        return best;
    }
};
PS_HD void proxyFromShape(DProxy& p, const Shape& s) {
    if (s.type == SHAPE_CIRCLE) {
        p.count = 1;
        p.vx = &s.px;
        p.vy = &s.py;
    } else {
        p.count = s.count;
        p.vx = s.vx;
        p.vy = s.vy;
    }
    p.radius = s.radius;
}

struct SimplexCache {
    float metric;
    int count;
    int indexA[3], indexB[3];
};
struct SimplexVertex {
    V2 wA, wB, w;
    float a;
    int indexA, indexB;
};
struct Simplex {
    SimplexVertex v[3];
    int count;
};
PS_HD float simplexMetric(const Simplex& s) {
    if (s.count == 2) return dist(s.v[0].w, s.v[1].w);
    if (s.count == 3) return cross(s.v[1].w - s.v[0].w, s.v[2].w - s.v[0].w);
    return 0.0f;
}
PS_HD void simplexSolve2(Simplex& s) {
    V2 w1 = s.v[0].w, w2 = s.v[1].w;
    V2 e12 = w2 - w1;
    float d12_2 = -dot(w1, e12);
    if (d12_2 <= 0.0f) {
        s.v[0].a = 1.0f;
        s.count = 1;
        return;
    }
    float d12_1 = dot(w2, e12);
    if (d12_1 <= 0.0f) {
        s.v[1].a = 1.0f;
        s.count = 1;
        s.v[0] = s.v[1];
        return;
    }
    float inv = 1.0f / (d12_1 + d12_2);
    s.v[0].a = d12_1 * inv;
    s.v[1].a = d12_2 * inv;
    s.count = 2;
}
PS_HD void simplexSolve3(Simplex& s) {
    V2 w1 = s.v[0].w, w2 = s.v[1].w, w3 = s.v[2].w;
    V2 e12 = w2 - w1;
    float w1e12 = dot(w1, e12), w2e12 = dot(w2, e12);
    float d12_1 = w2e12, d12_2 = -w1e12;
    V2 e13 = w3 - w1;
    float w1e13 = dot(w1, e13), w3e13 = dot(w3, e13);
    float d13_1 = w3e13, d13_2 = -w1e13;
    V2 e23 = w3 - w2;
    float w2e23 = dot(w2, e23), w3e23 = dot(w3, e23);
    float d23_1 = w3e23, d23_2 = -w2e23;
    float n123 = cross(e12, e13);
    float d123_1 = n123 * cross(w2, w3);
    float d123_2 = n123 * cross(w3, w1);
    float d123_3 = n123 * cross(w1, w2);
    if (d12_2 <= 0.0f && d13_2 <= 0.0f) {
        s.v[0].a = 1.0f;
        s.count = 1;
        return;
    }
    if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {
        float inv = 1.0f / (d12_1 + d12_2);
        s.v[0].a = d12_1 * inv;
        s.v[1].a = d12_2 * inv;
        s.count = 2;
        return;
    }
    if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {
        float inv = 1.0f / (d13_1 + d13_2);
        s.v[0].a = d13_1 * inv;
        s.v[2].a = d13_2 * inv;
        s.count = 2;
        s.v[1] = s.v[2];
        return;
    }
    if (d12_1 <= 0.0f && d23_2 <= 0.0f) {
        s.v[1].a = 1.0f;
        s.count = 1;
        s.v[0] = s.v[1];
        return;
    }
    if (d13_1 <= 0.0f && d23_1 <= 0.0f) {
        s.v[2].a = 1.0f;
        s.count = 1;
        s.v[0] = s.v[2];
        return;
    }
    if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {
        float inv = 1.0f / (d23_1 + d23_2);
        s.v[1].a = d23_1 * inv;
        s.v[2].a = d23_2 * inv;
        s.count = 2;
        s.v[0] = s.v[2];
        return;
    }
    float inv = 1.0f / (d123_1 + d123_2 + d123_3);
    s.v[0].a = d123_1 * inv;
    s.v[1].a = d123_2 * inv;
    s.v[2].a = d123_3 * inv;
    s.count = 3;
}

// Distance.distance with useRadii = false; returns the core distance and refreshes the cache
PS_HDN float gjkDistance(SimplexCache& cache, const DProxy& pA, const Xf& xfA, const DProxy& pB, const Xf& xfB) {
    Simplex sx;
    // Simplex.readCache
    sx.count = cache.count;
    for (int i = 0; i < sx.count; ++i) {
        SimplexVertex& sv = sx.v[i];
        sv.indexA = cache.indexA[i];
        sv.indexB = cache.indexB[i];
        sv.wA = mulX(xfA, pA.v(sv.indexA));
        sv.wB = mulX(xfB, pB.v(sv.indexB));
        sv.w = sv.wB - sv.wA;
        sv.a = 0.0f;
    }
    if (sx.count > 1) {
        float metric1 = cache.metric, metric2 = simplexMetric(sx);
        if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < kEpsilon) sx.count = 0;
    }
    if (sx.count == 0) {
        SimplexVertex& sv = sx.v[0];
        sv.indexA = 0;
        sv.indexB = 0;
        sv.wA = mulX(xfA, pA.v(0));
        sv.wB = mulX(xfB, pB.v(0));
        sv.w = sv.wB - sv.wA;
        sx.count = 1;
    }
    int saveA[3], saveB[3], saveCount = 0;
    int iter = 0;
    while (iter < 20) {
        saveCount = sx.count;
        for (int i = 0; i < saveCount; ++i) {
            saveA[i] = sx.v[i].indexA;
            saveB[i] = sx.v[i].indexB;
        }
        if (sx.count == 2) simplexSolve2(sx);
        else if (sx.count == 3) simplexSolve3(sx);
        if (sx.count == 3) break;
        // Simplex.getSearchDirection
        V2 d;
        if (sx.count == 1) d = -sx.v[0].w;
        else {
            V2 e12 = sx.v[1].w - sx.v[0].w;
            float sgn = cross(e12, -sx.v[0].w);
            d = sgn > 0.0f ? crossSV(1.0f, e12) : crossVS(e12, 1.0f);
        }
        if (dot(d, d) < kEpsilon * kEpsilon) break;
        SimplexVertex& nv = sx.v[sx.count];
        nv.indexA = pA.support(mulTR(xfA, -d));
        nv.wA = mulX(xfA, pA.v(nv.indexA));
        nv.indexB = pB.support(mulTR(xfB, d));
        nv.wB = mulX(xfB, pB.v(nv.indexB));
        nv.w = nv.wB - nv.wA;
        ++iter;
        bool duplicate = false;
        for (int i = 0; i < saveCount; ++i)
            if (nv.indexA == saveA[i] && nv.indexB == saveB[i]) {
                duplicate = true;
                break;
            }
        if (duplicate) break;
        ++sx.count;
    }
    // witness points
    V2 wpA = mk(0, 0), wpB = mk(0, 0);
    if (sx.count == 1) {
        wpA = sx.v[0].wA;
        wpB = sx.v[0].wB;
    } else if (sx.count == 2) {
        wpA = sx.v[0].a * sx.v[0].wA + sx.v[1].a * sx.v[1].wA;
        wpB = sx.v[0].a * sx.v[0].wB + sx.v[1].a * sx.v[1].wB;
    } else if (sx.count == 3) {
        wpA = sx.v[0].a * sx.v[0].wA + sx.v[1].a * sx.v[1].wA + sx.v[2].a * sx.v[2].wA;
        wpB = wpA;
    }
    float distance = dist(wpA, wpB);
    cache.metric = simplexMetric(sx);
    cache.count = sx.count;
    for (int i = 0; i < sx.count; ++i) {
        cache.indexA[i] = sx.v[i].indexA;
        cache.indexB[i] = sx.v[i].indexB;
    }
    return distance;
}

enum { TOI_UNKNOWN = 0, TOI_FAILED, TOI_OVERLAPPED, TOI_TOUCHING, TOI_SEPARATED };
enum { SEP_POINTS = 0, SEP_FACE_A = 1, SEP_FACE_B = 2 };

struct SepFn {
    int type;
    V2 localPoint, axis;
};
PS_HD void sepInit(SepFn& f, const SimplexCache& cache, const DProxy& pA, const Xf& xfA, const DProxy& pB, const Xf& xfB) {
    if (cache.count == 1) {
        f.type = SEP_POINTS;
        V2 pointA = mulX(xfA, pA.v(cache.indexA[0])), pointB = mulX(xfB, pB.v(cache.indexB[0]));
        f.axis = pointB - pointA;
        normalize(f.axis);
    } else if (cache.indexA[0] == cache.indexA[1]) {
        f.type = SEP_FACE_B;
        V2 b1 = pB.v(cache.indexB[0]), b2 = pB.v(cache.indexB[1]);
        f.axis = crossVS(b2 - b1, 1.0f);
        normalize(f.axis);
        V2 normal = mulR(xfB, f.axis);
        f.localPoint = 0.5f * (b1 + b2);
        V2 pointB = mulX(xfB, f.localPoint);
        V2 pointA = mulX(xfA, pA.v(cache.indexA[0]));
        float s = dot(pointA - pointB, normal);
        if (s < 0.0f) f.axis = -f.axis;
    } else {
        f.type = SEP_FACE_A;
        V2 a1 = pA.v(cache.indexA[0]), a2 = pA.v(cache.indexA[1]);
        f.axis = crossVS(a2 - a1, 1.0f);
        normalize(f.axis);
        V2 normal = mulR(xfA, f.axis);
        f.localPoint = 0.5f * (a1 + a2);
        V2 pointA = mulX(xfA, f.localPoint);
        V2 pointB = mulX(xfB, pB.v(cache.indexB[0]));
        float s = dot(pointB - pointA, normal);
        if (s < 0.0f) f.axis = -f.axis;
    }
}
PS_HD float sepFindMin(const SepFn& f, int& indexA, int& indexB, const DProxy& pA, const Xf& xfA, const DProxy& pB, const Xf& xfB) {
    if (f.type == SEP_POINTS) {
        V2 axisA = mulTR(xfA, f.axis), axisB = mulTR(xfB, -f.axis);
        indexA = pA.support(axisA);
        indexB = pB.support(axisB);
        V2 pointA = mulX(xfA, pA.v(indexA)), pointB = mulX(xfB, pB.v(indexB));
        return dot(pointB - pointA, f.axis);
    } else if (f.type == SEP_FACE_A) {
        V2 normal = mulR(xfA, f.axis);
        V2 pointA = mulX(xfA, f.localPoint);
        V2 axisB = mulTR(xfB, -normal);
        indexA = -1;
        indexB = pB.support(axisB);
        V2 pointB = mulX(xfB, pB.v(indexB));
        return dot(pointB - pointA, normal);
    } else {
        V2 normal = mulR(xfB, f.axis);
        V2 pointB = mulX(xfB, f.localPoint);
        V2 axisA = mulTR(xfA, -normal);
        indexB = -1;
        indexA = pA.support(axisA);
        V2 pointA = mulX(xfA, pA.v(indexA));
        return dot(pointA - pointB, normal);
    }
}
PS_HD float sepEvaluate(const SepFn& f, int indexA, int indexB, const DProxy& pA, const Xf& xfA, const DProxy& pB, const Xf& xfB) {
    if (f.type == SEP_POINTS) {
        V2 pointA = mulX(xfA, pA.v(indexA)), pointB = mulX(xfB, pB.v(indexB));
        return dot(pointB - pointA, f.axis);
    } else if (f.type == SEP_FACE_A) {
        V2 normal = mulR(xfA, f.axis);
        V2 pointA = mulX(xfA, f.localPoint);
        V2 pointB = mulX(xfB, pB.v(indexB));
        return dot(pointB - pointA, normal);
    } else {
        V2 normal = mulR(xfB, f.axis);
        V2 pointB = mulX(xfB, f.localPoint);
        V2 pointA = mulX(xfA, pA.v(indexA));
        return dot(pointA - pointB, normal);
    }
}

// TimeOfImpact.timeOfImpact; kindA / kindB select the exact transform shortcut of each side (see sweepXfKind).
PS_HDN void timeOfImpact(int& outState, float& outT, Trig tr, const DProxy& pA, Sweep sweepA, int kindA, const DProxy& pB, Sweep sweepB, int kindB,
                         const Xf& staticXf, float tMax) {
    outState = TOI_UNKNOWN;
    outT = tMax;
    sweepNormalize(sweepA);
    sweepNormalize(sweepB);
    float totalRadius = pA.radius + pB.radius;
    float target = fmaxf_(kLinearSlop, totalRadius - 3.0f * kLinearSlop);
    float tolerance = 0.25f * kLinearSlop;
    float t1 = 0.0f;
    int iter = 0;
    SimplexCache cache;
    cache.count = 0;
    cache.metric = 0.0f;
    for (;;) {
        // One outer iteration is a pure function of (t1, simplex cache). If it ends without a verdict and leaves both
        // unchanged (the push-back loop ran out of its 8 rounds), every later iteration repeats it bit for bit until the
        // iteration cap turns the result into FAILED at t1 -- so that verdict can be returned right away.
        float t1In = t1;
        SimplexCache cacheIn = cache;
        Xf xfA = sweepXfKind(tr, sweepA, kindA, staticXf, t1), xfB = sweepXfKind(tr, sweepB, kindB, staticXf, t1);
        float d = gjkDistance(cache, pA, xfA, pB, xfB);
        if (d <= 0.0f) {
            outState = TOI_OVERLAPPED;
            outT = 0.0f;
            break;
        }
        if (d < target + tolerance) {
            outState = TOI_TOUCHING;
            outT = t1;
            break;
        }
        SepFn fcn;
        sepInit(fcn, cache, pA, xfA, pB, xfB);
        bool done = false;
        float t2 = tMax;
        int pushBackIter = 0;
        for (;;) {
            int indexA, indexB;
            Xf xfA2 = sweepXfKind(tr, sweepA, kindA, staticXf, t2), xfB2 = sweepXfKind(tr, sweepB, kindB, staticXf, t2);
            float s2 = sepFindMin(fcn, indexA, indexB, pA, xfA2, pB, xfB2);
            if (s2 > target + tolerance) {
                outState = TOI_SEPARATED;
                outT = tMax;
                done = true;
                break;
            }
            if (s2 > target - tolerance) {
                t1 = t2;
                break;
            }
            Xf xfA1 = sweepXfKind(tr, sweepA, kindA, staticXf, t1), xfB1 = sweepXfKind(tr, sweepB, kindB, staticXf, t1);
            float s1 = sepEvaluate(fcn, indexA, indexB, pA, xfA1, pB, xfB1);
            if (s1 < target - tolerance) {
                outState = TOI_FAILED;
                outT = t1;
                done = true;
                break;
            }
            if (s1 <= target + tolerance) {
                outState = TOI_TOUCHING;
                outT = t1;
                done = true;
                break;
            }
            int rootIterCount = 0;
            float a1 = t1, a2 = t2;
            for (;;) {
                float t;
                if (rootIterCount & 1) t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
                else t = 0.5f * (a1 + a2);
                Xf xfAt = sweepXfKind(tr, sweepA, kindA, staticXf, t), xfBt = sweepXfKind(tr, sweepB, kindB, staticXf, t);
                float s = sepEvaluate(fcn, indexA, indexB, pA, xfAt, pB, xfBt);
                if (fabsf(s - target) < tolerance) {
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
            if (pushBackIter == 8) break;
        }
        ++iter;
        if (done) break;
        if (iter == 20) {
            outState = TOI_FAILED;
            outT = t1;
            break;
        }
        if (iter > 1 && t1 == t1In && cache.count == cacheIn.count && cache.metric == cacheIn.metric) {
            bool same = true;
            for (int i = 0; i < cache.count; ++i) same = same && cache.indexA[i] == cacheIn.indexA[i] && cache.indexB[i] == cacheIn.indexB[i];
            if (same) {
                outState = TOI_FAILED;
                outT = t1;
                break;
            }
        }
    }
}

}  // namespace ps
