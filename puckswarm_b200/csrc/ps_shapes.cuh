// puckswarm_b200 -- fixture shape tables of the PuckSwarm scene and their mass data.
//
// Scene: one static wall body of 84 two-vertex polygon "edges"
// (src/utils/FixtureUtils.java:51-74, src/arena/RoundedRectangleEnclosure.java:33-51), pucks of two
// concentric circles (src/arena/Puck.java:30-55, src/arena/DotPuck.java:22-34) and robots of 14 fixtures
// (src/arena/Robot.java:95-163). Shapes are built on the host once per engine with JBox2D's float
// arithmetic (PolygonShape.set / setAsEdge / computeMass, Body.resetMassData) and live in one table:
//   [0, 84) wall edges, [84, 86) puck outer / inner disc, [86, 100) robot fixtures in creation order.
#pragma once
#include "ps_math.cuh"

namespace ps {

static constexpr int kWallFixtures = 84;
static constexpr int kPuckFixtures = 2;
static constexpr int kRobotFixtures = 14;
static constexpr int kShapePuck0 = kWallFixtures;
static constexpr int kShapeRobot0 = kWallFixtures + kPuckFixtures;
static constexpr int kNumShapes = kShapeRobot0 + kRobotFixtures;
static constexpr int kWallGrid = 16;
static constexpr int kMaxVerts = 5;

enum { SHAPE_CIRCLE = 0, SHAPE_POLYGON = 1 };

struct Shape {
    int type, count;
    float radius;
    float px, py;          // circle centre (m_p)
    float cx, cy;          // polygon centroid
    float vx[kMaxVerts], vy[kMaxVerts], nx[kMaxVerts], ny[kMaxVerts];
    uint16_t category, mask;
    PS_HD V2 v(int i) const { return mk(vx[i], vy[i]); }
    PS_HD V2 n(int i) const { return mk(nx[i], ny[i]); }
};

struct BodyConst {      // Body.resetMassData results shared by every body of a kind
    float mass, invMass, I, invI, lcx, lcy;
};

// Everything a kernel needs to know about the static part of the scene.
struct Scene {
    Shape shapes[kNumShapes];
    Aabb wallFat[kWallFixtures];   // DynamicTree fat AABB of the wall proxies (never move)
    BodyConst puck, robot;
    Xf wallXf;                     // transform of the static wall body: position (0,0), R.set(0)
    float W, H, R;                 // enclosure WIDTH, HEIGHT, corner radius
    float cornerX, cornerY;        // |x|, |y| of the corner-arc centres
    float friction;                // sqrt(0.2 * 0.2): every fixture keeps FixtureDef's default friction
    int wallSensable;              // any wall fixture AABB within 22 cm of the origin (PointSensor body filter)
    uint32_t filt[kNumShapes];     // Filter of every shape: category | mask << 16
    // pair-search pruning: which wall proxies' fat boxes reach each cell of a kWallGrid x kWallGrid grid over the arena
    uint32_t wallGrid[kWallGrid * kWallGrid][3];
    float wgX0, wgY0, wgInvX, wgInvY;
};
PS_HD bool filtCollide(uint32_t a, uint32_t b) { return ((a >> 16) & (b & 0xFFFFu)) != 0 && ((a & 0xFFFFu) & (b >> 16)) != 0; }
PS_HD int wallGridIndex(float v, float v0, float inv) {
    int i = (int)floorf((v - v0) * inv);
    return i < 0 ? 0 : (i > kWallGrid - 1 ? kWallGrid - 1 : i);
}

// ---- proxy id <-> (body, shape) mapping; ids are fixture creation order ----
struct Dims {
    int P, R, NB, NP;              // pucks, robots, dynamic bodies, dynamic proxies
};
PS_HD int proxyBody(const Dims& d, int id) {   // -1 = wall
    if (id < kWallFixtures) return -1;
    int k = id - kWallFixtures;
    if (k < 2 * d.P) return k >> 1;
    return d.P + (k - 2 * d.P) / kRobotFixtures;
}
PS_HD int proxyShape(const Dims& d, int id) {
    if (id < kWallFixtures) return id;
    int k = id - kWallFixtures;
    if (k < 2 * d.P) return kShapePuck0 + (k & 1);
    return kShapeRobot0 + (k - 2 * d.P) % kRobotFixtures;
}
PS_HD int bodyFirstProxy(const Dims& d, int b) { return kWallFixtures + (b < d.P ? 2 * b : 2 * d.P + (b - d.P) * kRobotFixtures); }
PS_HD int bodyProxyCount(const Dims& d, int b) { return b < d.P ? kPuckFixtures : kRobotFixtures; }
PS_HD int bodyFirstShape(const Dims& d, int b) { return b < d.P ? kShapePuck0 : kShapeRobot0; }
// ContactManager.addPair filter: Filter categories / masks (group index is always 0 here)
PS_HD bool shouldCollide(const Shape& a, const Shape& b) { return (a.mask & b.category) != 0 && (a.category & b.mask) != 0; }

// Shape.computeAABB
PS_HD Aabb shapeAabb(const Shape& s, const Xf& xf) {
    Aabb r;
    if (s.type == SHAPE_CIRCLE) {
        V2 c = mk(xf.px, xf.py) + mulR(xf, mk(s.px, s.py));
        r.lx = c.x - s.radius;
        r.ly = c.y - s.radius;
        r.ux = c.x + s.radius;
        r.uy = c.y + s.radius;
    } else {
        V2 lo = mulX(xf, s.v(0)), up = lo;
        for (int i = 1; i < s.count; ++i) {
            V2 w = mulX(xf, s.v(i));
            lo = mk(fminf_(lo.x, w.x), fminf_(lo.y, w.y));
            up = mk(fmaxf_(up.x, w.x), fmaxf_(up.y, w.y));
        }
        r.lx = lo.x - s.radius;
        r.ly = lo.y - s.radius;
        r.ux = up.x + s.radius;
        r.uy = up.y + s.radius;
    }
    return r;
}

// Shape.testPoint (polygon radius ignored)
PS_HD bool shapeTestPoint(const Shape& s, const Xf& xf, V2 p) {
    if (s.type == SHAPE_CIRCLE) {
        V2 c = mk(xf.px, xf.py) + mulR(xf, mk(s.px, s.py));
        V2 d = p - c;
        return dot(d, d) <= s.radius * s.radius;
    }
    V2 pl = mulTR(xf, mk(p.x - xf.px, p.y - xf.py));
    for (int i = 0; i < s.count; ++i)
        if (dot(s.n(i), pl - s.v(i)) > 0.0f) return false;
    return true;
}

// ------------------------------------------------------------------ host-side construction
namespace build {

inline Shape circle(float r, float px, float py, uint16_t cat, uint16_t mask) {
    Shape s = {};
    s.type = SHAPE_CIRCLE;
    s.radius = r;
    s.px = px;
    s.py = py;
    s.category = cat;
    s.mask = mask;
    return s;
}
// PolygonShape.set: normals = normalize(cross(edge, 1)); centroid = area-weighted triangle fan about (0,0)
inline Shape polygon(const V2* vs, int n, uint16_t cat, uint16_t mask) {
    Shape s = {};
    s.type = SHAPE_POLYGON;
    s.count = n;
    s.radius = kPolygonRadius;
    s.category = cat;
    s.mask = mask;
    for (int i = 0; i < n; i++) {
        s.vx[i] = vs[i].x;
        s.vy[i] = vs[i].y;
    }
    for (int i = 0; i < n; i++) {
        V2 e = vs[(i + 1) % n] - vs[i];
        V2 nn = crossVS(e, 1.0f);
        normalize(nn);
        s.nx[i] = nn.x;
        s.ny[i] = nn.y;
    }
    V2 c = mk(0, 0);
    float area = 0.0f;
    const float inv3 = 1.0f / 3.0f;
    for (int i = 0; i < n; i++) {
        V2 p1 = mk(0, 0), p2 = vs[i], p3 = vs[(i + 1) % n];
        float D = cross(p2 - p1, p3 - p1);
        float tri = 0.5f * D;
        area += tri;
        V2 t = (tri * inv3) * (p1 + p2 + p3);
        c = c + t;
    }
    float ia = 1.0f / area;
    s.cx = c.x * ia;
    s.cy = c.y * ia;
    return s;
}
// PolygonShape.setAsEdge
inline Shape edge(V2 a, V2 b, uint16_t cat, uint16_t mask) {
    Shape s = {};
    s.type = SHAPE_POLYGON;
    s.count = 2;
    s.radius = kPolygonRadius;
    s.category = cat;
    s.mask = mask;
    s.vx[0] = a.x;
    s.vy[0] = a.y;
    s.vx[1] = b.x;
    s.vy[1] = b.y;
    V2 m = 0.5f * (a + b);
    s.cx = m.x;
    s.cy = m.y;
    V2 nn = crossVS(b - a, 1.0f);
    normalize(nn);
    s.nx[0] = nn.x;
    s.ny[0] = nn.y;
    s.nx[1] = -nn.x;
    s.ny[1] = -nn.y;
    return s;
}
// FixtureUtils.createArc: the running angle is a double, end points are cx + r * (float)cos(angle)
inline void arc(Shape* out, float cx, float cy, float r, float a0, float a1, int n, uint16_t cat, uint16_t mask) {
    float total = a1 - a0;
    double delta = total / n;
    double ang = a0;
    float x0 = cx + r * (float)cos(ang), y0 = cy + r * (float)sin(ang);
    for (int i = 0; i < n; i++) {
        ang += delta;
        float x1 = cx + r * (float)cos(ang), y1 = cy + r * (float)sin(ang);
        out[i] = edge(mk(x0, y0), mk(x1, y1), cat, mask);
        x0 = x1;
        y0 = y1;
    }
}
// Shape.computeMass -> (mass, centre, I about the body origin)
inline void massOf(const Shape& s, float density, float* mass, V2* centre, float* I) {
    if (s.type == SHAPE_CIRCLE) {
        *mass = density * kPi * s.radius * s.radius;
        *centre = mk(s.px, s.py);
        *I = *mass * (0.5f * s.radius * s.radius + (s.px * s.px + s.py * s.py));
        return;
    }
    if (s.count == 2) {
        *centre = 0.5f * (s.v(0) + s.v(1));
        *mass = 0.0f;
        *I = 0.0f;
        return;
    }
    V2 c = mk(0, 0);
    float area = 0, inertia = 0;
    const float k3 = 1.0f / 3.0f;
    for (int i = 0; i < s.count; i++) {
        V2 p1 = mk(0, 0), p2 = s.v(i), p3 = s.v((i + 1) % s.count);
        V2 e1 = p2 - p1, e2 = p3 - p1;
        float D = cross(e1, e2);
        float tri = 0.5f * D;
        area += tri;
        c = c + (tri * k3) * (p1 + p2 + p3);
        float px = p1.x, py = p1.y, ex1 = e1.x, ey1 = e1.y, ex2 = e2.x, ey2 = e2.y;
        float intx2 = k3 * (0.25f * (ex1 * ex1 + ex2 * ex1 + ex2 * ex2) + (px * ex1 + px * ex2)) + 0.5f * px * px;
        float inty2 = k3 * (0.25f * (ey1 * ey1 + ey2 * ey1 + ey2 * ey2) + (py * ey1 + py * ey2)) + 0.5f * py * py;
        inertia += D * (intx2 + inty2);
    }
    *mass = density * area;
    float ia = 1.0f / area;
    *centre = mk(c.x * ia, c.y * ia);
    *I = density * inertia;
}
// Body.resetMassData over fixtures first..first+n-1 (iterated newest first, like the fixture list)
inline BodyConst bodyMass(const Shape* shapes, int first, int n, float density) {
    float mass = 0, I = 0;
    V2 centre = mk(0, 0);
    for (int i = first + n - 1; i >= first; i--) {
        float m, inert;
        V2 c;
        massOf(shapes[i], density, &m, &c, &inert);
        mass += m;
        centre = centre + m * c;
        I += inert;
    }
    BodyConst b;
    b.invMass = 1.0f / mass;
    centre = mk(centre.x * b.invMass, centre.y * b.invMass);
    I -= mass * dot(centre, centre);
    b.mass = mass;
    b.I = I;
    b.invI = 1.0f / I;
    b.lcx = centre.x;
    b.lcy = centre.y;
    return b;
}

inline void scene(Scene* sc, float scale, const Trig& trig) {
    const float PIf = (float)3.14159265358979323846, PI_2f = (float)(3.14159265358979323846 / 2.0);
    const float TO_RADf = (float)(3.14159265358979323846 / 180);
    // --- enclosure (RoundedRectangleEnclosure.java:37-50, FixtureUtils.java:51-74); default filter 1 / 0xFFFF
    float W = scale * 187.0f, H = scale * 187.0f, R = scale * 15.5f;
    sc->W = W;
    sc->H = H;
    sc->R = R;
    sc->cornerX = W / 2 - R;
    sc->cornerY = H / 2 - R;
    float leftX = -W / 2, rightX = W / 2, topY = H / 2, bottomY = -H / 2;
    Shape* s = sc->shapes;
    arc(s + 0, leftX + R, topY - R, R, PI_2f, PIf, 20, 1, 0xFFFF);
    arc(s + 20, rightX - R, topY - R, R, 0, PI_2f, 20, 1, 0xFFFF);
    arc(s + 40, leftX + R, bottomY + R, R, PIf, 3 * PIf / 2, 20, 1, 0xFFFF);
    arc(s + 60, rightX - R, bottomY + R, R, 3 * PIf / 2, 2 * PIf, 20, 1, 0xFFFF);
    s[80] = edge(mk(leftX + R, topY), mk(rightX - R, topY), 1, 0xFFFF);
    s[81] = edge(mk(leftX + R, bottomY), mk(rightX - R, bottomY), 1, 0xFFFF);
    s[82] = edge(mk(leftX, bottomY + R), mk(leftX, topY - R), 1, 0xFFFF);
    s[83] = edge(mk(rightX, bottomY + R), mk(rightX, topY - R), 1, 0xFFFF);
    // --- puck (Puck.java:20,42-56; DotPuck.java:20-34): outer disc r = WIDTH/2, inner r = (2/3 WIDTH)/2, cat 4 / mask 5
    const float PUCK_WIDTH = 6.3f, INNER_WIDTH = 2 / 3.0f * PUCK_WIDTH;
    s[kShapePuck0 + 0] = circle(PUCK_WIDTH / 2, 0, 0, 0x04, 0x05);
    s[kShapePuck0 + 1] = circle(INNER_WIDTH / 2, 0, 0, 0x04, 0x05);
    // --- robot (Robot.java:95-163)
    float shiftX = -3.88f, rx = shiftX + 6.4f;
    V2 hull[5] = {mk(shiftX - 8.7f, 0.0f), mk(shiftX - 7.7f, -5.7f), mk(rx, -5.7f), mk(rx, 5.7f), mk(shiftX - 7.7f, 5.7f)};
    Shape* r = s + kShapeRobot0;
    r[0] = polygon(hull, 5, 0x01, 0x07);
    float innerY = 4.5f;
    V2 top[3] = {mk(rx, innerY), mk(shiftX + 11.5f, 3.29f), mk(rx, 5.7f)};
    V2 bot[3] = {mk(rx, -5.7f), mk(shiftX + 11.5f, -3.29f), mk(rx, -innerY)};
    r[1] = polygon(top, 3, 0x01, 0x07);
    r[2] = polygon(bot, 3, 0x01, 0x07);
    float startDeg = 65.0f, stopDeg = 360.0f - startDeg;
    arc(r + 3, shiftX + 10.0f, 0.0f, 3.6f, startDeg * TO_RADf, stopDeg * TO_RADf, 10, 0x01, 0x07);
    r[13] = circle(3.6f, shiftX + 10.0f, 0.0f, 0x02, 0x03);
    sc->puck = bodyMass(s, kShapePuck0, kPuckFixtures, 5.0f);
    sc->robot = bodyMass(s, kShapeRobot0, kRobotFixtures, 5.0f);
    sc->friction = sqrtf(0.2f * 0.2f);
    // wall proxies: tight AABB at the identity transform, fattened by aabbExtension (Fixture.createProxy)
    Xf id;
    id.px = 0.0f;
    id.py = 0.0f;
    id.c = psCos(trig, 0.0f);
    id.s = psSin(trig, 0.0f);
    sc->wallXf = id;
    sc->wallSensable = 0;
    for (int i = 0; i < kWallFixtures; i++) {
        Aabb a = shapeAabb(s[i], id);
        // PointSensor only reaches the wall body when the pixel is within 22 cm of the body origin (0, 0)
        float nx = a.lx > 0 ? a.lx : (a.ux < 0 ? a.ux : 0), ny = a.ly > 0 ? a.ly : (a.uy < 0 ? a.uy : 0);
        if (nx * nx + ny * ny <= 22.0f * 22.0f * 1.01f) sc->wallSensable = 1;
        a.lx -= kAabbExtension;
        a.ly -= kAabbExtension;
        a.ux += kAabbExtension;
        a.uy += kAabbExtension;
        sc->wallFat[i] = a;
    }
    for (int i = 0; i < kNumShapes; i++) sc->filt[i] = (uint32_t)s[i].category | ((uint32_t)s[i].mask << 16);
    // wall grid over [-W/2 - 1, W/2 + 1] x [-H/2 - 1, H/2 + 1]; a cell lists every wall box within 0.05 of it (the margin
    // absorbs the rounding of the cell-index computation)
    float gx0 = -W / 2 - 1.0f, gy0 = -H / 2 - 1.0f, csx = (W + 2.0f) / kWallGrid, csy = (H + 2.0f) / kWallGrid;
    sc->wgX0 = gx0;
    sc->wgY0 = gy0;
    sc->wgInvX = 1.0f / csx;
    sc->wgInvY = 1.0f / csy;
    for (int ix = 0; ix < kWallGrid; ix++)
        for (int iy = 0; iy < kWallGrid; iy++) {
            Aabb c;
            c.lx = gx0 + ix * csx - 0.05f;
            c.ux = gx0 + (ix + 1) * csx + 0.05f;
            c.ly = gy0 + iy * csy - 0.05f;
            c.uy = gy0 + (iy + 1) * csy + 0.05f;
            uint32_t* m = sc->wallGrid[ix * kWallGrid + iy];
            m[0] = m[1] = m[2] = 0;
            for (int i = 0; i < kWallFixtures; i++)
                if (aabbOverlap(sc->wallFat[i], c)) m[i >> 5] |= 1u << (i & 31);
        }
}

}  // namespace build

}  // namespace ps
