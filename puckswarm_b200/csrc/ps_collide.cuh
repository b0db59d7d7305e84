// puckswarm_b200 -- narrow phase: circle-circle, polygon-circle and polygon-polygon manifolds with
// JBox2D 2.1.2 / Box2D v2.1.2 semantics (org.jbox2d.collision.Collision).
// Polygons here have 2..5 vertices (two-vertex "edges" from PolygonShape.setAsEdge included).
#pragma once
#include "ps_shapes.cuh"

namespace ps {

enum { MF_CIRCLES = 0, MF_FACE_A = 1, MF_FACE_B = 2 };

// ContactID packed in 8 bits: referenceEdge (3) | incidentEdge (3) << 3 | incidentVertex << 6 | flip << 7
PS_HD uint32_t packId(int refEdge, int incEdge, int incVertex, int flip) {
    return (uint32_t)(refEdge | (incEdge << 3) | (incVertex << 6) | (flip << 7));
}

struct Manifold {
    int type, pointCount;
    V2 localNormal, localPoint;
    V2 p[2];          // ManifoldPoint.localPoint
    uint32_t id[2];   // packed ContactID
};

PS_HD void collideCircles(Manifold& m, const Shape& A, const Xf& xfA, const Shape& B, const Xf& xfB) {
    m.pointCount = 0;
    V2 pA = mulX(xfA, mk(A.px, A.py)), pB = mulX(xfB, mk(B.px, B.py));
    V2 d = pB - pA;
    float distSqr = dot(d, d);
    float radius = A.radius + B.radius;
    if (distSqr > radius * radius) return;
    m.type = MF_CIRCLES;
    m.localPoint = mk(A.px, A.py);
    m.localNormal = mk(0, 0);
    m.pointCount = 1;
    m.p[0] = mk(B.px, B.py);
    m.id[0] = 0;
}

PS_HD void collidePolygonAndCircle(Manifold& m, const Shape& A, const Xf& xfA, const Shape& B, const Xf& xfB) {
    m.pointCount = 0;
    V2 c = mulX(xfB, mk(B.px, B.py));
    V2 cLocal = mulTX(xfA, c);
    int normalIndex = 0;
    float separation = -kMaxFloat;
    float radius = A.radius + B.radius;
    int n = A.count;
    for (int i = 0; i < n; ++i) {
        float s = dot(A.n(i), cLocal - A.v(i));
        if (s > radius) return;
        if (s > separation) {
            separation = s;
            normalIndex = i;
        }
    }
    int i1 = normalIndex, i2 = i1 + 1 < n ? i1 + 1 : 0;
    V2 v1 = A.v(i1), v2 = A.v(i2);
    m.type = MF_FACE_A;
    m.p[0] = mk(B.px, B.py);
    m.id[0] = 0;
    if (separation < kEpsilon) {
        m.pointCount = 1;
        m.localNormal = A.n(normalIndex);
        m.localPoint = 0.5f * (v1 + v2);
        return;
    }
    float u1 = dot(cLocal - v1, v2 - v1);
    float u2 = dot(cLocal - v2, v1 - v2);
    if (u1 <= 0.0f) {
        if (distSq(cLocal, v1) > radius * radius) return;
        m.pointCount = 1;
        m.localNormal = cLocal - v1;
        normalize(m.localNormal);
        m.localPoint = v1;
    } else if (u2 <= 0.0f) {
        if (distSq(cLocal, v2) > radius * radius) return;
        m.pointCount = 1;
        m.localNormal = cLocal - v2;
        normalize(m.localNormal);
        m.localPoint = v2;
    } else {
        V2 faceCenter = 0.5f * (v1 + v2);
        float sep = dot(cLocal - faceCenter, A.n(i1));
        if (sep > radius) return;
        m.pointCount = 1;
        m.localNormal = A.n(i1);
        m.localPoint = faceCenter;
    }
}

PS_HD float edgeSeparation(const Shape& p1, const Xf& xf1, int edge1, const Shape& p2, const Xf& xf2) {
    V2 n1w = mulR(xf1, p1.n(edge1));
    V2 n1 = mulTR(xf2, n1w);
    int index = 0;
    float minDot = kMaxFloat;
    for (int i = 0; i < p2.count; ++i) {
        float d = dot(p2.v(i), n1);
        if (d < minDot) {
            minDot = d;
            index = i;
        }
    }
    V2 v1 = mulX(xf1, p1.v(edge1));
    V2 v2 = mulX(xf2, p2.v(index));
    return dot(v2 - v1, n1w);
}

PS_HD float findMaxSeparation(int& edgeIndex, const Shape& p1, const Xf& xf1, const Shape& p2, const Xf& xf2) {
    int count1 = p1.count;
    V2 d = mulX(xf2, mk(p2.cx, p2.cy)) - mulX(xf1, mk(p1.cx, p1.cy));
    V2 dLocal1 = mulTR(xf1, d);
    int edge = 0;
    float maxDot = -kMaxFloat;
    for (int i = 0; i < count1; ++i) {
        float dt = dot(p1.n(i), dLocal1);
        if (dt > maxDot) {
            maxDot = dt;
            edge = i;
        }
    }
    float s = edgeSeparation(p1, xf1, edge, p2, xf2);
    int prevEdge = edge - 1 >= 0 ? edge - 1 : count1 - 1;
    float sPrev = edgeSeparation(p1, xf1, prevEdge, p2, xf2);
    int nextEdge = edge + 1 < count1 ? edge + 1 : 0;
    float sNext = edgeSeparation(p1, xf1, nextEdge, p2, xf2);
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
        edgeIndex = edge;
        return s;
    }
    for (;;) {
        if (increment == -1) edge = bestEdge - 1 >= 0 ? bestEdge - 1 : count1 - 1;
        else edge = bestEdge + 1 < count1 ? bestEdge + 1 : 0;
        s = edgeSeparation(p1, xf1, edge, p2, xf2);
        if (s > bestSeparation) {
            bestEdge = edge;
            bestSeparation = s;
        } else
            break;
    }
    edgeIndex = bestEdge;
    return bestSeparation;
}

struct ClipVertex {
    V2 v;
    uint32_t id;
};
// Collision.clipSegmentToLine
PS_HD int clipSegment(ClipVertex* out, const ClipVertex* in, V2 normal, float offset) {
    int numOut = 0;
    float d0 = dot(normal, in[0].v) - offset;
    float d1 = dot(normal, in[1].v) - offset;
    if (d0 <= 0.0f) out[numOut++] = in[0];
    if (d1 <= 0.0f) out[numOut++] = in[1];
    if (d0 * d1 < 0.0f) {
        float interp = d0 / (d0 - d1);
        out[numOut].v = in[0].v + interp * (in[1].v - in[0].v);
        out[numOut].id = d0 > 0.0f ? in[0].id : in[1].id;
        ++numOut;
    }
    return numOut;
}

PS_HDN void collidePolygons(Manifold& m, const Shape& polyA, const Xf& xfA, const Shape& polyB, const Xf& xfB) {
    m.pointCount = 0;
    float totalRadius = polyA.radius + polyB.radius;
    int edgeA = 0;
    float separationA = findMaxSeparation(edgeA, polyA, xfA, polyB, xfB);
    if (separationA > totalRadius) return;
    int edgeB = 0;
    float separationB = findMaxSeparation(edgeB, polyB, xfB, polyA, xfA);
    if (separationB > totalRadius) return;
    const Shape *poly1, *poly2;
    Xf xf1, xf2;
    int edge1, flip;
    if (separationB > 0.98f * separationA + 0.001f) {
        poly1 = &polyB;
        poly2 = &polyA;
        xf1 = xfB;
        xf2 = xfA;
        edge1 = edgeB;
        m.type = MF_FACE_B;
        flip = 1;
    } else {
        poly1 = &polyA;
        poly2 = &polyB;
        xf1 = xfA;
        xf2 = xfB;
        edge1 = edgeA;
        m.type = MF_FACE_A;
        flip = 0;
    }
    // findIncidentEdge
    ClipVertex incident[2];
    {
        V2 n1 = mulTR(xf2, mulR(xf1, poly1->n(edge1)));
        int index = 0;
        float minDot = kMaxFloat;
        for (int i = 0; i < poly2->count; ++i) {
            float d = dot(n1, poly2->n(i));
            if (d < minDot) {
                minDot = d;
                index = i;
            }
        }
        int i1 = index, i2 = i1 + 1 < poly2->count ? i1 + 1 : 0;
        incident[0].v = mulX(xf2, poly2->v(i1));
        incident[0].id = packId(edge1, i1, 0, 0);
        incident[1].v = mulX(xf2, poly2->v(i2));
        incident[1].id = packId(edge1, i2, 1, 0);
    }
    int count1 = poly1->count;
    V2 v11 = poly1->v(edge1);
    V2 v12 = edge1 + 1 < count1 ? poly1->v(edge1 + 1) : poly1->v(0);
    V2 localTangent = v12 - v11;
    normalize(localTangent);
    V2 localNormal = crossVS(localTangent, 1.0f);
    V2 planePoint = 0.5f * (v11 + v12);
    V2 tangent = mulR(xf1, localTangent);
    V2 normal = crossVS(tangent, 1.0f);
    v11 = mulX(xf1, v11);
    v12 = mulX(xf1, v12);
    float frontOffset = dot(normal, v11);
    float sideOffset1 = -dot(tangent, v11) + totalRadius;
    float sideOffset2 = dot(tangent, v12) + totalRadius;
    ClipVertex clip1[2], clip2[2];
    int np = clipSegment(clip1, incident, -tangent, sideOffset1);
    if (np < 2) return;
    np = clipSegment(clip2, clip1, tangent, sideOffset2);
    if (np < 2) return;
    m.localNormal = localNormal;
    m.localPoint = planePoint;
    int pointCount = 0;
    for (int i = 0; i < 2; ++i) {
        float separation = dot(normal, clip2[i].v) - frontOffset;
        if (separation <= totalRadius) {
            m.p[pointCount] = mulTX(xf2, clip2[i].v);
            m.id[pointCount] = (clip2[i].id & 0x7Fu) | ((uint32_t)flip << 7);
            ++pointCount;
        }
    }
    m.pointCount = pointCount;
}

// Contact.evaluate: dispatch by shape types (polygon is always A in polygon-circle contacts)
PS_HD void evaluateManifold(Manifold& m, const Shape& A, const Xf& xfA, const Shape& B, const Xf& xfB) {
    if (A.type == SHAPE_CIRCLE && B.type == SHAPE_CIRCLE) collideCircles(m, A, xfA, B, xfB);
    else if (A.type == SHAPE_POLYGON && B.type == SHAPE_CIRCLE) collidePolygonAndCircle(m, A, xfA, B, xfB);
    else collidePolygons(m, A, xfA, B, xfB);
}

// WorldManifold.initialize: world normal and the (at most 2) world contact points
PS_HD void worldManifold(const Manifold& m, const Xf& xfA, float radiusA, const Xf& xfB, float radiusB, V2& normal, V2* points) {
    if (m.type == MF_CIRCLES) {
        V2 pointA = mulX(xfA, m.localPoint);
        V2 pointB = mulX(xfB, m.p[0]);
        V2 nrm = mk(1.0f, 0.0f);
        if (distSq(pointA, pointB) > kEpsilon * kEpsilon) {
            nrm = pointB - pointA;
            normalize(nrm);
        }
        normal = nrm;
        V2 cA = pointA + radiusA * nrm;
        V2 cB = pointB - radiusB * nrm;
        points[0] = 0.5f * (cA + cB);
    } else if (m.type == MF_FACE_A) {
        V2 nrm = mulR(xfA, m.localNormal);
        V2 planePoint = mulX(xfA, m.localPoint);
        normal = nrm;
        for (int i = 0; i < m.pointCount; ++i) {
            V2 clipPoint = mulX(xfB, m.p[i]);
            V2 cA = clipPoint + (radiusA - dot(clipPoint - planePoint, nrm)) * nrm;
            V2 cB = clipPoint - radiusB * nrm;
            points[i] = 0.5f * (cA + cB);
        }
    } else {
        V2 nrm = mulR(xfB, m.localNormal);
        V2 planePoint = mulX(xfB, m.localPoint);
        for (int i = 0; i < m.pointCount; ++i) {
            V2 clipPoint = mulX(xfA, m.p[i]);
            V2 cB = clipPoint + (radiusB - dot(clipPoint - planePoint, nrm)) * nrm;
            V2 cA = clipPoint - radiusA * nrm;
            points[i] = 0.5f * (cA + cB);
        }
        normal = -nrm;
    }
}

// PositionSolverManifold.initialize for one manifold point
PS_HD void positionManifold(int type, V2 localNormal, V2 localPoint, V2 pointLocal, float radius, const Xf& xfA, const Xf& xfB,
                            V2& normal, V2& point, float& separation) {
    if (type == MF_CIRCLES) {
        V2 pointA = mulX(xfA, localPoint);
        V2 pointB = mulX(xfB, pointLocal);
        if (distSq(pointA, pointB) > kEpsilon * kEpsilon) {
            normal = pointB - pointA;
            normalize(normal);
        } else {
            normal = mk(1.0f, 0.0f);
        }
        point = 0.5f * (pointA + pointB);
        separation = dot(pointB - pointA, normal) - radius;
    } else if (type == MF_FACE_A) {
        normal = mulR(xfA, localNormal);
        V2 planePoint = mulX(xfA, localPoint);
        V2 clipPoint = mulX(xfB, pointLocal);
        separation = dot(clipPoint - planePoint, normal) - radius;
        point = clipPoint;
    } else {
        normal = mulR(xfB, localNormal);
        V2 planePoint = mulX(xfB, localPoint);
        V2 clipPoint = mulX(xfA, pointLocal);
        separation = dot(clipPoint - planePoint, normal) - radius;
        point = clipPoint;
        normal = -normal;
    }
}

}  // namespace ps
