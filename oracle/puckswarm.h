// ORACLE (test infrastructure only -- never linked into or called by the product path).
//
// CPU restatement of PuckSwarm's per-tick path above JBox2D: scene construction, spawn, Arena.step /
// Arena.coast, the grid camera, LocalMap, BlobFinder, VFH+, and the CacheCons and BHD controllers.
// Each function cites the reference file:line it follows. PARITY IS UNPINNED: the reference has no
// tests or golden vectors and cannot be executed here (no JVM).
#pragma once
#include <limits>
#include "jb2d.h"
#include "jrandom.h"
#include "../include/puckswarm_b200.h"
#include <memory>
#include <cstdio>

namespace oracle {
using jb::Vec2;

// ---- utils.AngleUtils (src/utils/AngleUtils.java:5-42) ----
static const float PIf = (float)M_PI;
static const float PI_OVER_TWOf = (float)(M_PI / 2.0);
static const float TO_RADf = (float)(M_PI / 180);
static const float PI_OVER_4f = (float)(M_PI / 4.0);
inline double constrainAngle(double angle) {
    while (angle > M_PI) angle -= 2 * M_PI;
    while (angle <= -M_PI) angle += 2 * M_PI;
    return angle;
}
inline double getAngularDifference(double a, double b) {
    a = constrainAngle(a);
    b = constrainAngle(b);
    double error = std::fabs(a - b);
    if (error > M_PI) {
        error -= M_PI * 2;
        error = std::fabs(error);
    }
    return error;
}

// ---- constants of the scene ----
static const float SRV_LENGTH = 22.0f;                         // Robot.java:41
static const float MAX_FORWARDS = 5 * 150000;                  // Robot.java:44
static const float MAX_TORQUE = 5 * 475000;                    // Robot.java:46
static const int ONE_EIGHTY_TIME = 10;                         // Robot.java:49
static const float ROBOT_MIN_BOUNDARY_DISTANCE = SRV_LENGTH / 2;  // Robot.java:52
static const float PUCK_WIDTH = 6.3f;                          // Puck.java:20
static const float PUCK_MIN_BOUNDARY_DISTANCE = PUCK_WIDTH / 2;   // Puck.java:24
static const float INNER_WIDTH = 2 / 3.0f * PUCK_WIDTH;        // DotPuck.java:20
static const float RADIUS_OF_LARGEST_OBJECT_SQD = SRV_LENGTH * SRV_LENGTH;  // PointSensor.java:40

// ---- utils.FixtureUtils (src/utils/FixtureUtils.java:22-74) ----
inline void createArc(jb::World& w, jb::Body* body, float cx, float cy, float radius, float angle0, float angle1, int n, jb::FixtureDef def) {
    float totalAngle = angle1 - angle0;
    double angleDelta = totalAngle / n;
    double angle = angle0;
    float x0 = cx + radius * (float)std::cos(angle);
    float y0 = cy + radius * (float)std::sin(angle);
    for (int i = 0; i < n; i++) {
        angle += angleDelta;
        float x1 = cx + radius * (float)std::cos(angle);
        float y1 = cy + radius * (float)std::sin(angle);
        def.shape = jb::Shape::edge(Vec2(x0, y0), Vec2(x1, y1));
        w.createFixture(body, def);
        x0 = x1;
        y0 = y1;
    }
}

// ---- arena.RoundedRectangleEnclosure (src/arena/RoundedRectangleEnclosure.java:33-93) ----
struct Enclosure {
    float WIDTH, HEIGHT, R;
    Vec2 upperLeft, upperRight, lowerLeft, lowerRight;
    jb::Body* body = nullptr;
    JavaRandom* rng = nullptr;

    void build(jb::World& w, float scale, JavaRandom* rng_) {
        rng = rng_;
        jb::BodyDef bd;
        body = w.createBody(bd);
        WIDTH = scale * 187.0f;
        HEIGHT = scale * 187.0f;
        R = scale * 15.5f;
        upperLeft = Vec2(-WIDTH / 2 + R, HEIGHT / 2 - R);
        upperRight = Vec2(WIDTH / 2 - R, HEIGHT / 2 - R);
        lowerLeft = Vec2(-WIDTH / 2 + R, -HEIGHT / 2 + R);
        lowerRight = Vec2(WIDTH / 2 - R, -HEIGHT / 2 + R);
        float leftX = -WIDTH / 2, rightX = WIDTH / 2, topY = HEIGHT / 2, bottomY = -HEIGHT / 2, radius = R;
        const int nSegments = 20;
        jb::FixtureDef def;  // new FixtureDef(): density 0, friction 0.2, userData null
        def.userKind = jb::U_WALL;
        createArc(w, body, leftX + radius, topY - radius, radius, PI_OVER_TWOf, PIf, nSegments, def);
        createArc(w, body, rightX - radius, topY - radius, radius, 0, PI_OVER_TWOf, nSegments, def);
        createArc(w, body, leftX + radius, bottomY + radius, radius, PIf, 3 * PIf / 2, nSegments, def);
        createArc(w, body, rightX - radius, bottomY + radius, radius, 3 * PIf / 2, 2 * PIf, nSegments, def);
        def.shape = jb::Shape::edge(Vec2(leftX + radius, topY), Vec2(rightX - radius, topY));
        w.createFixture(body, def);
        def.shape = jb::Shape::edge(Vec2(leftX + radius, bottomY), Vec2(rightX - radius, bottomY));
        w.createFixture(body, def);
        def.shape = jb::Shape::edge(Vec2(leftX, bottomY + radius), Vec2(leftX, topY - radius));
        w.createFixture(body, def);
        def.shape = jb::Shape::edge(Vec2(rightX, bottomY + radius), Vec2(rightX, topY - radius));
        w.createFixture(body, def);
    }
    bool inFreeSpace(const Vec2& v, float minDistance) const {
        float ax = std::fabs(v.x), ay = std::fabs(v.y);
        if (ax < WIDTH / 2 - R - minDistance && ay < HEIGHT / 2 - minDistance) return true;
        else if (ax > WIDTH / 2 - R && ax < WIDTH / 2 - minDistance && ay < HEIGHT / 2 - R - minDistance) return true;
        else if (jb::distance(v, upperLeft) < R - minDistance) return true;
        else if (jb::distance(v, upperRight) < R - minDistance) return true;
        else if (jb::distance(v, lowerLeft) < R - minDistance) return true;
        else if (jb::distance(v, lowerRight) < R - minDistance) return true;
        return false;
    }
    bool getRandomInside(float minDistance, Vec2* out) {
        for (int i = 0; i < 100000; i++) {
            float fx = rng->nextFloat();
            float fy = rng->nextFloat();
            Vec2 v(fx * WIDTH - WIDTH / 2.0f, fy * HEIGHT - HEIGHT / 2.0f);
            if (inFreeSpace(v, minDistance)) {
                *out = v;
                return true;
            }
        }
        return false;
    }
};

// ---- sensors.GridCalibration (src/sensors/GridCalibration.java:53-114) ----
struct CalibPixel {
    bool valid = false;
    float Xr = 0, Yr = 0;
    bool gripperBody = false, gripperHole = false;
};
struct Calibration {
    int imageWidth = 0, imageHeight = 0, nHolePixels = 0;
    std::vector<CalibPixel> data;  // [x * imageHeight + y]
    const CalibPixel* get(int x, int y) const {
        const CalibPixel& c = data[(size_t)x * imageHeight + y];
        return c.valid ? &c : nullptr;
    }
    void load(const float* XrYr, const uint16_t* xy, const uint8_t* flags, int n, int w, int h) {
        imageWidth = w;
        imageHeight = h;
        data.assign((size_t)w * h, CalibPixel());
        nHolePixels = 0;
        for (int k = 0; k < n; k++) {
            CalibPixel& c = data[(size_t)xy[2 * k] * h + xy[2 * k + 1]];
            c.valid = true;
            c.Xr = XrYr[2 * k];
            c.Yr = XrYr[2 * k + 1];
            c.gripperBody = (flags[k] & 1) != 0;
            c.gripperHole = (flags[k] & 2) != 0;
        }
        for (const CalibPixel& c : data)
            if (c.valid && c.gripperHole) nHolePixels++;
    }
};

struct Cluster {  // localmap.Cluster (src/localmap/Cluster.java:12-57); subgraph kept as vertex list + degrees
    Vec2 centroid;
    int size = 0;
    int puckType = 0;
    std::vector<Vec2> verts;
    std::vector<int> degree;
    int id = -1;  // identity (index in rawClusters), -1 for ad-hoc one-puck clusters
    static Cluster single(Vec2 c, int k) {
        Cluster cl;
        cl.centroid = c;
        cl.size = 1;
        cl.puckType = k;
        cl.verts.push_back(c);
        cl.degree.push_back(0);
        return cl;
    }
};
inline bool vecEquals(const Vec2& a, const Vec2& b) {  // Vec2.equals: Float.floatToIntBits comparison
    return std::memcmp(&a.x, &b.x, 4) == 0 && std::memcmp(&a.y, &b.y, 4) == 0;
}

struct Blob {  // sensors.Blob (src/sensors/Blob.java:8-18)
    int x0, x1, y0, y1, area, cx, cy;
};

// ---- localmap.LocalMap (src/localmap/LocalMap.java) ----
struct LocalMap {
    const Calibration* calib = nullptr;
    int width = 0, height = 0;
    float minXr, maxXr, minYr, maxYr;
    std::vector<uint8_t> occupancy;  // [i * height + j]
    std::vector<Vec2> pucks[PS_MAX_COLOURS];
    bool carrying = false;
    Vec2 carriedV;
    int carriedType = PS_T_NOTHING;
    int carriedCluster = -1;  // id into rawClusters
    std::vector<Cluster> rawClusters;
    std::vector<int> clusters;  // filtered list: ids into rawClusters
    float CLUSTER_DISTANCE_THRESHOLD = 1.5f, CARRY_THRESHOLD_LO = 0.1f, CARRY_THRESHOLD_HI = 0.4f;
    static constexpr float ROBOT_THRESHOLD_DISTANCE_SQD = 100.0f;

    void init(const Calibration* c) {  // :90-121
        calib = c;
        minXr = INFINITY;
        maxXr = -INFINITY;
        minYr = INFINITY;
        maxYr = -INFINITY;
        for (int i = 0; i < c->imageWidth; i++)
            for (int j = 0; j < c->imageHeight; j++) {
                const CalibPixel* p = c->get(i, j);
                if (p) {
                    minXr = std::min(minXr, p->Xr);
                    maxXr = std::max(maxXr, p->Xr);
                    minYr = std::min(minYr, p->Yr);
                    maxYr = std::max(maxYr, p->Yr);
                }
            }
        width = (int)(0.5 + (maxYr - minYr) / 1.0f) + 1;
        height = (int)(0.5 + (maxXr) / 1.0f) + 1;
        occupancy.assign((size_t)width * height, PS_T_NOTHING);
    }
    void getGridPoint(float Xr, float Yr, int* i, int* j) const {  // :443-447
        *i = (int)(0.5 + (maxYr - Yr) / 1.0f);
        *j = (int)(0.5 + (Xr) / 1.0f);
    }
    Vec2 getGroundPlane(int i, int j) const { return Vec2(1.0f * j, maxYr - 1.0f * i); }  // :452-454
    uint8_t& occ(int i, int j) { return occupancy[(size_t)i * height + j]; }
    uint8_t occ(int i, int j) const { return occupancy[(size_t)i * height + j]; }

    void update(const std::vector<uint8_t>& image /* [x*H + y] */) {  // :183-208
        extractOccupancy(image);
        extractPucks(image);
        rawClusters.clear();
        for (int k = 0; k < PS_MAX_COLOURS; k++) extractClusters(pucks[k], k, rawClusters, CLUSTER_DISTANCE_THRESHOLD);
        for (size_t i = 0; i < rawClusters.size(); i++) rawClusters[i].id = (int)i;
        carriedCluster = -1;
        if (carrying) {
            for (const Cluster& cl : rawClusters)
                for (const Vec2& v : cl.verts)
                    if (vecEquals(carriedV, v)) carriedCluster = cl.id;
        }
        clusters.clear();
        for (size_t i = 0; i < rawClusters.size(); i++) clusters.push_back((int)i);
        filterClusters();
    }
    void extractOccupancy(const std::vector<uint8_t>& image) {  // :210-229
        std::fill(occupancy.begin(), occupancy.end(), (uint8_t)PS_T_NOTHING);
        int H = calib->imageHeight;
        for (int i = 0; i < calib->imageWidth; i++)
            for (int j = 0; j < H; j++) {
                const CalibPixel* p = calib->get(i, j);
                if (p && !p->gripperHole) {
                    uint8_t type = image[(size_t)i * H + j];
                    if (type == PS_T_HIDDEN) continue;
                    int gi, gj;
                    getGridPoint(p->Xr, p->Yr, &gi, &gj);
                    occ(gi, gj) = type;
                }
            }
    }
    void extractPucks(const std::vector<uint8_t>& image);  // :236-317 (below, needs BlobFinder)
    static void extractClusters(const std::vector<Vec2>& pucks, int k, std::vector<Cluster>& out, float threshold) {  // :326-370
        // SimpleGraph vertex set: insertion-ordered, duplicates (Vec2.equals) dropped
        std::vector<Vec2> verts;
        for (const Vec2& v : pucks) {
            bool dup = false;
            for (const Vec2& u : verts)
                if (vecEquals(u, v)) dup = true;
            if (!dup) verts.push_back(v);
        }
        int n = (int)verts.size();
        std::vector<std::vector<int>> adj(n);
        for (int a = 0; a < n; a++)
            for (int b = 0; b < n; b++)
                if (!vecEquals(verts[a], verts[b]))
                    if (jb::distance(verts[a], verts[b]) < threshold) {
                        if (std::find(adj[a].begin(), adj[a].end(), b) == adj[a].end()) {
                            adj[a].push_back(b);
                            adj[b].push_back(a);
                        }
                    }
        // ConnectivityInspector.connectedSets: BFS from each unseen vertex in insertion order
        std::vector<int> comp(n, -1);
        int nc = 0;
        for (int s = 0; s < n; s++) {
            if (comp[s] != -1) continue;
            std::vector<int> queue{s};
            comp[s] = nc;
            for (size_t q = 0; q < queue.size(); q++)
                for (int o : adj[queue[q]])
                    if (comp[o] == -1) {
                        comp[o] = nc;
                        queue.push_back(o);
                    }
            Cluster cl;
            cl.puckType = k;
            // members in base-graph insertion order (UndirectedSubgraph.vertexSet). The reference sums the centroid
            // in HashSet order; for sizes <= 2 (always, at the committed 1.5 cm threshold) the sum is order-free.
            Vec2 centroid;
            for (int a = 0; a < n; a++)
                if (comp[a] == nc) {
                    cl.verts.push_back(verts[a]);
                    cl.degree.push_back((int)adj[a].size());
                    centroid += verts[a];
                }
            cl.size = (int)cl.verts.size();
            centroid *= 1.0f / cl.size;
            cl.centroid = centroid;
            out.push_back(cl);
            nc++;
        }
    }
    bool isReachable(const Vec2& v) const {  // :548-558
        double circleYr = 12.5, circleXr = 0, radius = circleYr + 0;
        double dXr = v.x - circleXr;
        double dYr = std::min(std::fabs(circleYr - v.y), std::fabs(circleYr + v.y));
        return std::sqrt(dXr * dXr + dYr * dYr) > radius;
    }
    bool isNeighbour(const Cluster& a, const Cluster& b) const {  // Cluster.java:49-57
        for (const Vec2& p : a.verts)
            for (const Vec2& q : b.verts)
                if (jb::distance(p, q) < CLUSTER_DISTANCE_THRESHOLD) return true;
        return false;
    }
    void filterClusters() {  // :375-437 (the centroid-distance filter is disabled: MAX_CLUSTER_DISTANCE_SQD = Float.MAX_VALUE)
        for (int i = 0; i < width; i++)
            for (int j = 0; j < height; j++)
                if (occ(i, j) == PS_T_ROBOT) {
                    Vec2 robotV = getGroundPlane(i, j);
                    for (size_t c = 0; c < clusters.size();) {
                        bool removed = false;
                        for (const Vec2& puckV : rawClusters[clusters[c]].verts)
                            if (jb::distanceSquared(robotV, puckV) < ROBOT_THRESHOLD_DISTANCE_SQD) {
                                clusters.erase(clusters.begin() + c);
                                removed = true;
                                break;
                            }
                        if (!removed) c++;
                    }
                }
        for (size_t c = 0; c < clusters.size();) {
            const Cluster& cluster = rawClusters[clusters[c]];
            bool remove = false;
            if (carrying && cluster.id == carriedCluster) remove = true;
            else if (!isReachable(cluster.centroid)) remove = true;
            else if (carrying && carriedType != cluster.puckType) remove = true;
            else if (carrying) {
                for (int o : clusters) {
                    const Cluster& other = rawClusters[o];
                    if (cluster.id != other.id && cluster.puckType != other.puckType && isNeighbour(cluster, other)) {
                        remove = true;
                        break;
                    }
                }
            }
            if (remove) clusters.erase(clusters.begin() + c);
            else c++;
        }
    }
    int getFreeerSide(int* leftOut = nullptr, int* rightOut = nullptr) const {  // :565-587
        int w_2 = width / 2, leftSum = 0, rightSum = 0;
        for (int i = 0; i < w_2; i++)
            for (int j = 0; j < height; j++)
                if (occ(i, j) == PS_T_ROBOT || occ(i, j) == PS_T_WALL) leftSum++;
        for (int i = w_2; i < width; i++)
            for (int j = 0; j < height; j++)
                if (occ(i, j) == PS_T_ROBOT || occ(i, j) == PS_T_WALL) rightSum++;
        if (leftOut) *leftOut = leftSum;
        if (rightOut) *rightOut = rightSum;
        return leftSum <= rightSum ? 1 : -1;
    }
    bool getClosestPuckToPositionAsCluster(const Vec2& v, float thr, Cluster* out) const {  // :481-506
        bool have = false;
        Vec2 closest;
        int closestType = -1;
        float smallest = INFINITY;
        for (int c : clusters)
            for (const Vec2& puckV : rawClusters[c].verts) {
                float ds = jb::distanceSquared(v, puckV);
                if (ds < thr && ds < smallest) {
                    smallest = ds;
                    closest = puckV;
                    closestType = rawClusters[c].puckType;
                    have = true;
                }
            }
        if (have) *out = Cluster::single(closest, closestType);
        return have;
    }
    bool getClosestClusterToPosition(const Vec2& v, float thr, Cluster* out) const {  // :513-528
        bool have = false;
        float smallest = INFINITY;
        for (int c : clusters) {
            if (rawClusters[c].id == carriedCluster) continue;
            float ds = jb::distanceSquared(v, rawClusters[c].centroid);
            if (ds < thr && ds < smallest) {
                smallest = ds;
                *out = rawClusters[c];
                have = true;
            }
        }
        return have;
    }
    void postFilterClusters(int puckType) {  // :600-607
        for (size_t c = 0; c < clusters.size();)
            if (rawClusters[clusters[c]].puckType == puckType) clusters.erase(clusters.begin() + c);
            else c++;
    }
    void postFilterOccupancy(int type) {  // :612-617
        for (auto& o : occupancy)
            if (o == type) o = PS_T_NOTHING;
    }
};

// ---- sensors.BlobFinder (src/sensors/BlobFinder.java:30-101) ----
struct BlobFinder {
    const std::vector<uint8_t>& image;
    int W, H;
    std::vector<float> labels;  // arena.Grid of floats, as in the reference
    int xMin, xMax, yMin, yMax, type;
    int blobArea, blobX0, blobX1, blobY0, blobY1;
    float label;
    BlobFinder(const std::vector<uint8_t>& img, int w, int h, int type_, int xMin_, int xMax_, int yMin_, int yMax_)
        : image(img), W(w), H(h), labels((size_t)w * h), xMin(xMin_), xMax(xMax_), yMin(yMin_), yMax(yMax_), type(type_) {}
    std::vector<Blob> getBlobs() {
        std::fill(labels.begin(), labels.end(), 0.0f);
        std::vector<Blob> blobs;
        label = 1;
        for (int i = xMin; i <= xMax; i++)
            for (int j = yMin; j < yMax; j++)
                if (image[(size_t)i * H + j] == type && labels[(size_t)i * H + j] == 0) {
                    blobArea = 0;
                    blobX0 = INT32_MAX;
                    blobX1 = -INT32_MAX;
                    blobY0 = INT32_MAX;
                    blobY1 = -INT32_MAX;
                    recursiveBlobber(i, j);
                    Blob b;
                    b.x0 = blobX0;
                    b.x1 = blobX1;
                    b.y0 = blobY0;
                    b.y1 = blobY1;
                    b.area = blobArea;
                    b.cx = (b.x0 + b.x1) / 2;
                    b.cy = (b.y0 + b.y1) / 2;
                    blobs.push_back(b);
                    label++;
                }
        return blobs;
    }
    void recursiveBlobber(int i, int j) {
        if (i < xMin || i > xMax || j < yMin || j > yMax) return;
        if (image[(size_t)i * H + j] != type || labels[(size_t)i * H + j] != 0) return;
        blobArea++;
        if (i < blobX0) blobX0 = i;
        if (i > blobX1) blobX1 = i;
        if (j < blobY0) blobY0 = j;
        if (j > blobY1) blobY1 = j;
        labels[(size_t)i * H + j] = label;
        recursiveBlobber(i - 1, j - 1);
        recursiveBlobber(i, j - 1);
        recursiveBlobber(i + 1, j - 1);
        recursiveBlobber(i - 1, j);
        recursiveBlobber(i + 1, j);
        recursiveBlobber(i - 1, j + 1);
        recursiveBlobber(i, j + 1);
        recursiveBlobber(i + 1, j + 1);
    }
};

inline void LocalMap::extractPucks(const std::vector<uint8_t>& image) {  // LocalMap.java:236-317
    bool lastCarrying = carrying;
    carrying = false;
    carriedType = PS_T_NOTHING;
    int W = calib->imageWidth, H = calib->imageHeight;
    for (int k = 0; k < PS_MAX_COLOURS; k++) {
        pucks[k].clear();
        BlobFinder finder(image, W, H, k, 0, W - 1, 0, H - 1);
        std::vector<Blob> blobs = finder.getBlobs();
        for (size_t b = 0; b < blobs.size();) {
            if (!calib->get(blobs[b].cx, blobs[b].cy)) blobs.erase(blobs.begin() + b);
            else b++;
        }
        int holeBlob = -1;
        for (size_t b = 0; b < blobs.size(); b++)
            if (calib->get(blobs[b].cx, blobs[b].cy)->gripperHole)
                if (holeBlob == -1 || blobs[b].area > blobs[holeBlob].area) holeBlob = (int)b;
        for (size_t b = 0; b < blobs.size();) {
            if ((int)b != holeBlob && calib->get(blobs[b].cx, blobs[b].cy)->gripperHole) {
                blobs.erase(blobs.begin() + b);
                if ((int)b < holeBlob) holeBlob--;
            } else
                b++;
        }
        if (holeBlob != -1) {
            const CalibPixel* holeCalib = calib->get(blobs[holeBlob].cx, blobs[holeBlob].cy);
            for (size_t b = 0; b < blobs.size();) {
                if ((int)b == holeBlob) {
                    b++;
                    continue;
                }
                const CalibPixel* thisCalib = calib->get(blobs[b].cx, blobs[b].cy);
                float dx = thisCalib->Xr - holeCalib->Xr;
                float dy = thisCalib->Yr - holeCalib->Yr;
                if (std::sqrt((double)(dx * dx + dy * dy)) < INNER_WIDTH) {
                    blobs.erase(blobs.begin() + b);
                    if ((int)b < holeBlob) holeBlob--;
                } else
                    b++;
            }
            float holeFraction = blobs[holeBlob].area / (float)calib->nHolePixels;
            if ((lastCarrying && holeFraction > CARRY_THRESHOLD_LO) || (!carrying && holeFraction > CARRY_THRESHOLD_HI)) {
                carrying = true;
                carriedType = k;
                carriedV = Vec2(holeCalib->Xr, holeCalib->Yr);
            }
        }
        for (const Blob& b : blobs) {
            const CalibPixel* p = calib->get(b.cx, b.cy);
            pucks[k].push_back(Vec2(p->Xr, p->Yr));
        }
    }
}

// ---- localmap.MovementCommand / MovementUtils (MovementCommand.java:20-23, MovementUtils.java:13-58) ----
inline float sat(float x, float a, float b) { return x <= a ? a : (x >= b ? b : x); }
inline float getSpeedScale(float desiredTurn) {
    float speedScale = (PI_OVER_4f - std::fabs(desiredTurn)) / PI_OVER_4f;
    return sat(speedScale, 0.25f, 1);
}
inline float getTorqueScale(float desiredTurn) {
    float K = 2.0f;
    return sat((float)(K * desiredTurn / (0.5 * M_PI)), -1, 1);
}
struct MovementCommand {
    float forwards = 0, torque = 0;
    MovementCommand() {}
    MovementCommand(float forwardSpeed, float turnAngle) {
        forwards = MAX_FORWARDS * forwardSpeed * getSpeedScale(turnAngle);
        torque = MAX_TORQUE * getTorqueScale(turnAngle);
    }
};

// ---- localmap.VFHPlus (src/localmap/VFHPlus.java) ----
struct VFHPlus {
    static const int N = 72 + 1;
    static constexpr float START_ANGLE = (float)(M_PI / 2);
    static constexpr float STOP_ANGLE = (float)(-M_PI / 2);
    static constexpr float ALPHA = (START_ANGLE - STOP_ANGLE) / (N - 1);
    static constexpr float MAX_SENSED_DISTANCE_SQD = 10000.0f / 3;  // (float) Math.pow(100, 2) / 3
    static constexpr float A = 1.0f;
    static constexpr float B = 1.0f / MAX_SENSED_DISTANCE_SQD;
    static const int S_MAX = 8;
    float MU_1 = 5, MU_2 = 2, MU_3 = 3;  // goalDirected = true in both shipped controllers
    float SAFETY_DISTANCE_GAMMA = 2, SAFETY_DISTANCE_MASK = 5, TAU_LO = 0.7f, TAU_HI = 0.85f;
    bool goalDirected = true;
    std::vector<float> baseMagnitude, beta, gamma;
    float primaryHistogram[N], binaryHistogram[N], maskedHistogram[N];
    struct CandidateDir {
        int sector;
        float cost;
    };
    std::vector<CandidateDir> dirs;
    int lastTurnSector = 0;
    float lastTurnAngle = 0;
    bool ignorePucks = false;
    bool inited = false;

    static float indexToAngle(int k) { return START_ANGLE - k * ALPHA; }
    static int javaIntCast(float f) {  // Java (int) of a float: NaN -> 0, saturating
        if (std::isnan(f)) return 0;
        if (f >= 2147483648.0f) return INT32_MAX;
        if (f <= -2147483648.0f) return INT32_MIN;
        return (int)f;
    }
    static int angleToIndex(float angle) { return javaIntCast((START_ANGLE - angle) / ALPHA); }

    void initDataStructures(const LocalMap& lm) {  // :170-217
        int w = lm.width, h = lm.height;
        beta.assign((size_t)w * h, 0);
        gamma.assign((size_t)w * h, 0);
        baseMagnitude.assign((size_t)w * h, 0);
        for (int i = 0; i < w; i++)
            for (int j = 0; j < h; j++) {
                Vec2 v = lm.getGroundPlane(i, j);
                float dSqd = v.x * v.x + v.y * v.y;
                baseMagnitude[(size_t)i * h + j] = std::max(0.0f, A - B * dSqd);
                beta[(size_t)i * h + j] = (float)std::atan2((double)v.y, (double)v.x);
                gamma[(size_t)i * h + j] = (float)(std::asin((double)SAFETY_DISTANCE_GAMMA / std::sqrt((double)dSqd)));
            }
        for (int k = 0; k < N; k++) primaryHistogram[k] = binaryHistogram[k] = maskedHistogram[k] = 0;
    }
    static bool isPuckType(uint8_t t) { return t < PS_MAX_COLOURS; }
    // returns false for Java null
    bool computeTurnAngle(const LocalMap& lm, float targetAngle, bool ignorePucks_, float* out) {  // :226-248
        if (!inited) {  // propertiesUpdated is set by the constructor, so the first call re-initialises
            initDataStructures(lm);
            inited = true;
        }
        ignorePucks = ignorePucks_;
        computePrimary(lm);
        computeBinary();
        computeMasked(lm);
        computeCandidateDirections(targetAngle);
        if (dirs.empty()) return false;
        const CandidateDir* best = &dirs[0];
        for (const CandidateDir& d : dirs)
            if (d.cost < best->cost) best = &d;
        lastTurnSector = best->sector;
        lastTurnAngle = indexToAngle(best->sector);
        *out = lastTurnAngle;
        return true;
    }
    void computePrimary(const LocalMap& lm) {  // :313-345
        for (int i = 0; i < N; i++) primaryHistogram[i] = 0;
        int w = lm.width, h = lm.height;
        for (int i = 0; i < w; i++)
            for (int j = 0; j < h; j++) {
                uint8_t t = lm.occ(i, j);
                if (t == PS_T_NOTHING || (ignorePucks && isPuckType(t))) continue;
                size_t c = (size_t)i * h + j;
                int startK = angleToIndex(beta[c] + gamma[c]);
                int stopK = angleToIndex(beta[c] - gamma[c]) + 1;
                for (int k = startK; k <= stopK; k++)
                    if (k >= 0 && k < N) primaryHistogram[k] = std::max(primaryHistogram[k], baseMagnitude[c]);
            }
    }
    void computeBinary() {  // :347-356
        for (int k = 0; k < N; k++) {
            if (primaryHistogram[k] > TAU_HI) binaryHistogram[k] = 1;
            else if (primaryHistogram[k] < TAU_LO) binaryHistogram[k] = 0;
        }
    }
    void computeMasked(const LocalMap& lm) {  // :358-403
        Vec2 leftCenter(-6.5f, 12.5f);
        Vec2 rightCenter(leftCenter.x, -leftCenter.y);
        float safetyRadiusSqd = (float)std::pow((double)(leftCenter.y + SAFETY_DISTANCE_MASK), 2.0);
        float leftLimit = START_ANGLE, rightLimit = STOP_ANGLE;
        int w = lm.width, h = lm.height;
        for (int i = 0; i < w; i++)
            for (int j = 0; j < h; j++) {
                uint8_t t = lm.occ(i, j);
                if (t == PS_T_NOTHING || (ignorePucks && isPuckType(t))) continue;
                Vec2 v = lm.getGroundPlane(i, j);
                float b = beta[(size_t)i * h + j];
                if (b < 0 && b > rightLimit && jb::distanceSquared(v, rightCenter) < safetyRadiusSqd) rightLimit = b;
                if (b > 0 && b < leftLimit && jb::distanceSquared(v, leftCenter) < safetyRadiusSqd) leftLimit = b;
            }
        for (int k = 0; k < N; k++) {
            maskedHistogram[k] = 0;
            if (binaryHistogram[k] == 0) {
                float sectorAngle = indexToAngle(k);
                if (sectorAngle > leftLimit || sectorAngle < rightLimit) maskedHistogram[k] = 1;
            } else
                maskedHistogram[k] = 1;
        }
    }
    void addCandidateDirs(int left, int right, int targetSector) {  // :440-466
        int width = right - left + 1;
        if (width < S_MAX) dirs.push_back({(left + right) / 2, 0});
        else {
            dirs.push_back({left + S_MAX / 2, 0});
            dirs.push_back({right - S_MAX / 2, 0});
            if (targetSector >= left && targetSector <= right) dirs.push_back({targetSector, 0});
        }
    }
    void computeCandidateDirections(float targetAngle) {  // :405-436
        int targetSector = angleToIndex(targetAngle);
        dirs.clear();
        bool opening = maskedHistogram[0] == 0;
        int left = 0;
        for (int k = 1; k < N; k++) {
            if (opening) {
                if (maskedHistogram[k - 1] == 0 && maskedHistogram[k] == 1) {
                    addCandidateDirs(left, k - 1, targetSector);
                    opening = false;
                }
            } else {
                if (maskedHistogram[k - 1] == 1 && maskedHistogram[k] == 0) {
                    left = k;
                    opening = true;
                }
            }
        }
        if (opening && maskedHistogram[N - 1] == 0) addCandidateDirs(left, N - 1, targetSector);
        for (CandidateDir& d : dirs) {
            double sectorAngle = indexToAngle(d.sector);
            d.cost = (float)(MU_1 * getAngularDifference(sectorAngle, targetAngle) + MU_2 * getAngularDifference(sectorAngle, 0) +
                             MU_3 * getAngularDifference(sectorAngle, lastTurnAngle));
        }
    }
};
inline MovementCommand applyVFH(VFHPlus& vfh, LocalMap& lm, float desiredTurn, bool ignorePucks) {  // MovementUtils.java:46-58
    float result;
    if (!vfh.computeTurnAngle(lm, desiredTurn, ignorePucks, &result)) return MovementCommand(0.5f, (float)(0.5f * M_PI * lm.getFreeerSide()));
    return MovementCommand(1, (float)(0.5f * M_PI * result));
}

// ---- sensors.Pose (src/sensors/Pose.java:31-55) ----
struct Pose {
    double x = 0, y = 0, theta = 0;
    static double getDistanceBetween(const Pose& a, const Pose& b) {
        double dx = b.x - a.x, dy = b.y - a.y;
        return std::sqrt(dx * dx + dy * dy);
    }
    static double getAngleFromTo(const Pose& a, const Pose& b) {
        double dx = b.x - a.x, dy = b.y - a.y;
        return constrainAngle(std::atan2(dy, dx) - a.theta);
    }
    static Pose getTranslated(const Pose& a, float xt, float yt) {
        double dist = std::sqrt((double)(xt * xt + yt * yt));
        double angle = std::atan2((double)yt, (double)xt) + a.theta;
        Pose p;
        p.x = a.x + dist * std::cos(angle);
        p.y = a.y + dist * std::sin(angle);
        p.theta = a.theta;
        return p;
    }
};

struct Params {  // the property keys of SURVEY C.8, carried in ps_config
    ps_config cfg;
};

// ---- controllers ----
struct Controller {
    virtual ~Controller() {}
    virtual void computeDesired(LocalMap& lm, const Pose& robotPose, int stepCount) = 0;
    MovementCommand movement;
    int state = 0, startCount = 0;
    int stateCounts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // what the controller wrote at the latest storage step: _R<i>_stateCounts.txt (FileUtils.saveArray of recentStateCounts,
    // CacheConsController.java:313-333 / BHDController.java:103-123) and _R<i>_cachePoints.txt (8 lines, NaN for null, :531-551)
    int savedStateCounts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float savedCache[PS_MAX_COLOURS][2] = {};
    int savedStep = -1;
    VFHPlus vfh;
    JavaRandom* random = nullptr;
    void updateStateCounts(int stepCount) {  // CacheConsController.java:313-333: the file output becomes savedStateCounts
        stateCounts[state]++;
        if (stepCount % 100 == 0) {
            for (int i = 0; i < 8; i++) savedStateCounts[i] = stateCounts[i];
            savedStep = stepCount;
            for (int i = 0; i < 8; i++) stateCounts[i] = 0;
        }
    }
};

// Deneubourg.java:13-24
inline float pickupProb(float f, float K1) {
    float p = (K1 / (K1 + f));
    return p * p;
}
inline float depositProb(float f, float K2) {
    float p = (f / (K2 + f));
    return p * p;
}

struct CacheConsController : Controller {  // src/controllers/CacheConsController.java
    const ps_config* P;
    bool haveTarget = false;
    Cluster target;
    bool cacheValid[PS_MAX_COLOURS];
    Pose cachePoints[PS_MAX_COLOURS];
    float rememberedCacheSizes[PS_MAX_COLOURS];
    int lastCarriedType = -1;
    int stepCount = 0;
    Pose robotPose;
    bool presetCaches = false;

    CacheConsController(const ps_config* cfg, JavaRandom* rng, bool preset) : P(cfg) {
        random = rng;
        presetCaches = preset;
        for (int k = 0; k < PS_MAX_COLOURS; k++) {
            cacheValid[k] = false;
            rememberedCacheSizes[k] = 0;
        }
        vfh.SAFETY_DISTANCE_GAMMA = cfg->vfh_safety_gamma;
        vfh.SAFETY_DISTANCE_MASK = cfg->vfh_safety_mask;
        vfh.TAU_LO = cfg->vfh_tau_lo * VFHPlus::A;
        vfh.TAU_HI = cfg->vfh_tau_hi * VFHPlus::A;
        if (preset) {  // :209-243
            float arenaWidth = 187.0f;
            int nTypes = 2;
            float L = 0.75f * arenaWidth / 2.0f;
            for (int type = 0; type < nTypes; type++) {
                float angle = (float)(type * (2 * M_PI) / nTypes + M_PI / 4.0f);
                float x = (float)(L * std::cos((double)angle));
                float y = (float)(L * std::sin((double)angle));
                cachePoints[type].x = x;
                cachePoints[type].y = y;
                cachePoints[type].theta = 0;
                cacheValid[type] = true;
                rememberedCacheSizes[type] = 1000;
            }
        }
    }
    bool ignorePucks() const { return presetCaches ? false : P->cc_ignore_pucks != 0; }
    void transition(int s) {
        state = s;
        startCount = stepCount;
    }
    Vec2 getCachePointInRobotCoords(int puckType) const {  // :695-705
        const Pose& cp = cachePoints[puckType];
        double d = Pose::getDistanceBetween(robotPose, cp);
        double alpha = Pose::getAngleFromTo(robotPose, cp);
        return Vec2((float)(d * (float)std::cos(alpha)), (float)(d * (float)std::sin(alpha)));
    }
    bool cachePointInCluster(const LocalMap& lm, const Cluster& c) const {  // :658-672
        Vec2 cp = getCachePointInRobotCoords(c.puckType);
        for (const Vec2& v : c.verts)
            if (jb::distance(v, cp) < lm.CLUSTER_DISTANCE_THRESHOLD) return true;
        return false;
    }
    bool carriedPuckInCacheCluster(const LocalMap& lm) const {  // :674-693
        if (lm.carriedCluster < 0) return false;
        Vec2 cp = getCachePointInRobotCoords(lm.carriedType);
        for (const Vec2& v : lm.rawClusters[lm.carriedCluster].verts)
            if (jb::distance(v, cp) < lm.CLUSTER_DISTANCE_THRESHOLD) return true;
        return false;
    }
    void checkForget() {  // :556-566
        if (P->cc_cache_selection == PS_SEL_RELATIVE_FORGET) {
            if (!presetCaches && P->cc_forget_prob != -1 && random->nextFloat() < P->cc_forget_prob)
                for (int k = 0; k < PS_MAX_COLOURS; k++) {
                    cacheValid[k] = false;
                    rememberedCacheSizes[k] = 0;
                }
        }
    }
    bool checkCacheSelection(const Cluster& cand) {  // :569-590
        switch (P->cc_cache_selection) {
            case PS_SEL_PROB_ABSOLUTE:
                if (random->nextFloat() < depositProb((float)cand.size, P->cc_k_absolute)) return true;
                break;
            case PS_SEL_PROB_RELATIVE: {
                float diff = cand.size - rememberedCacheSizes[cand.puckType];
                if (random->nextFloat() < depositProb(diff, P->cc_k_relative)) return true;
            } break;
            default:
                if (cand.size > rememberedCacheSizes[cand.puckType]) return true;
                break;
        }
        return false;
    }
    bool checkCacheSeparation(const Cluster& cand) {  // :598-647
        std::vector<int> nearby;
        for (int otherK = 0; otherK < PS_MAX_COLOURS; otherK++)
            if (otherK != cand.puckType && cacheValid[otherK]) {
                Vec2 other = getCachePointInRobotCoords(otherK);
                if (jb::distance(cand.centroid, other) < P->cc_home_separation_distance) nearby.push_back(otherK);
            }
        switch (P->cc_cache_separation) {
            case PS_SEP_IGNORE: return true;
            case PS_SEP_NULLIFY_OLD:
                for (int k : nearby) {
                    cacheValid[k] = false;
                    rememberedCacheSizes[k] = 0;
                }
                return true;
            case PS_SEP_ACCEPT_ISOLATED: return nearby.empty();
            default: {
                bool newOneIsLargest = true;
                for (int k : nearby)
                    if (cand.size <= rememberedCacheSizes[k]) newOneIsLargest = false;
                if (newOneIsLargest)
                    for (int k : nearby) {
                        cacheValid[k] = false;
                        rememberedCacheSizes[k] = 0;
                    }
                return newOneIsLargest;
            }
        }
    }
    // ExtremaSelector.selectCluster (ExtremaSelector.java:36-59); ALWAYS_TARGET_ISOLATED is false (C.7)
    bool selectCluster(const LocalMap& lm, bool preferLargest, Cluster* out) {
        if (lm.clusters.empty()) return false;
        const Cluster* extreme = &lm.rawClusters[lm.clusters[0]];
        for (int ci : lm.clusters) {
            const Cluster* cluster = &lm.rawClusters[ci];
            if (preferLargest && cluster->size > extreme->size) extreme = cluster;
            else if (!preferLargest && cluster->size < extreme->size) extreme = cluster;
            else if (cluster->size == extreme->size && std::fabs(cluster->centroid.y) < std::fabs(extreme->centroid.y)) extreme = cluster;
        }
        float prob;
        if (preferLargest) prob = P->cc_k2 < 0 ? 1.0f : depositProb((float)extreme->size, P->cc_k2);
        else prob = P->cc_k1 < 0 ? 1.0f : pickupProb((float)extreme->size, P->cc_k1);
        if (random->nextFloat() < prob) {
            *out = *extreme;
            return true;
        }
        return false;
    }
    // ClusterTargetSelector.selectTarget / selectPickupPuckCluster (ClusterTargetSelector.java:48-101)
    bool selectTarget(const LocalMap& lm, bool preferLargest, bool isolatePuck, Cluster* out) {
        Cluster tc;
        if (!selectCluster(lm, preferLargest, &tc)) return false;
        if (!isolatePuck) {
            *out = tc;
            return true;
        }
        float largestYr = -INFINITY, smallestYr = INFINITY;
        int li = -1, si = -1;
        for (size_t i = 0; i < tc.verts.size(); i++) {
            if (tc.verts[i].y > largestYr) {
                largestYr = tc.verts[i].y;
                li = (int)i;
            }
            if (tc.verts[i].y < smallestYr) {
                smallestYr = tc.verts[i].y;
                si = (int)i;
            }
        }
        Cluster one = tc.degree[li] <= tc.degree[si] ? Cluster::single(tc.verts[li], tc.puckType) : Cluster::single(tc.verts[si], tc.puckType);
        if (lm.isReachable(one.centroid)) {
            *out = one;
            return true;
        }
        return false;
    }
    // ProbSeekController.matchTarget (ProbSeekController.java:286-301)
    static bool matchTarget(const LocalMap& lm, const Cluster& last, float thr, Cluster* out) {
        Cluster nt;
        bool have = lm.carrying ? lm.getClosestClusterToPosition(last.centroid, thr, &nt) : lm.getClosestPuckToPositionAsCluster(last.centroid, thr, &nt);
        if (!have || !lm.isReachable(nt.centroid)) return false;
        *out = nt;
        return true;
    }
    void computeDesired(LocalMap& lm, const Pose& pose, int stepCount_) override {  // :336-554
        robotPose = pose;
        bool carrying = lm.carrying;
        int carriedType = lm.carriedType;
        stepCount = stepCount_;
        checkForget();
        for (int k = 0; k < PS_MAX_COLOURS; k++) {
            if (cacheValid[k]) {
                const Cluster* cacheCluster = nullptr;
                for (const Cluster& c : lm.rawClusters)
                    if (c.puckType == k && cachePointInCluster(lm, c)) cacheCluster = &c;
                if (cacheCluster) rememberedCacheSizes[k] = std::max(rememberedCacheSizes[k], (float)cacheCluster->size);
            }
        }
        for (int k = 0; k < PS_MAX_COLOURS; k++) {
            const Cluster* candidate = nullptr;
            for (const Cluster& c : lm.rawClusters)
                if (c.puckType == k && (candidate == nullptr || c.size > candidate->size)) candidate = &c;
            if (!candidate) continue;
            if (!checkCacheSelection(*candidate)) continue;
            if (checkCacheSeparation(*candidate)) {  // initializeCache :649-656
                int type = candidate->puckType;
                cachePoints[type] = Pose::getTranslated(robotPose, candidate->centroid.x, candidate->centroid.y);
                cacheValid[type] = true;
                rememberedCacheSizes[type] = (float)candidate->size;
            }
        }
        for (int k = 0; k < PS_MAX_COLOURS; k++)
            if (!cacheValid[k]) lm.postFilterClusters(k);

        if (carrying && !cacheValid[carriedType]) transition(PS_CC_SPIT_OUT);
        else {
            switch (state) {
                case PS_CC_PU_SCAN:
                    if (carrying) transition(PS_CC_HOMING);
                    else {
                        haveTarget = selectTarget(lm, carrying, !carrying, &target);
                        if (haveTarget) {
                            if (!cacheValid[target.puckType]) haveTarget = false;
                            else if (cachePointInCluster(lm, target)) haveTarget = false;
                            else transition(PS_CC_PU_TARGET);
                        }
                    }
                    break;
                case PS_CC_PU_TARGET:
                    if (carrying) transition(PS_CC_HOMING);
                    else {
                        Cluster nt;
                        haveTarget = matchTarget(lm, target, P->cc_prediction_distance_sqd, &nt);
                        if (haveTarget) target = nt;
                        if (!haveTarget) transition(PS_CC_PU_SCAN);
                        else if (!cacheValid[target.puckType]) transition(PS_CC_PU_SCAN);
                        else if (cachePointInCluster(lm, target)) transition(PS_CC_PU_SCAN);
                        else if (stepCount - startCount > P->cc_pick_up_time) transition(PS_CC_PU_SCAN);
                    }
                    break;
                case PS_CC_HOMING:
                    if (!carrying) transition(PS_CC_PU_SCAN);
                    else if (!cacheValid[carriedType]) transition(PS_CC_PU_SCAN);
                    else if (carriedPuckInCacheCluster(lm)) transition(PS_CC_DE_PUSH);
                    break;
                case PS_CC_DE_PUSH:
                    if (stepCount - startCount >= P->cc_push_in_time) transition(PS_CC_DE_BACKUP);
                    break;
                case PS_CC_DE_BACKUP:
                    if (stepCount - startCount >= P->cc_backup_time) {
                        if (!cacheValid[lastCarriedType]) transition(PS_CC_PU_SCAN);
                        else transition(PS_CC_EXILE);
                    }
                    break;
                case PS_CC_EXILE:
                    if (stepCount - startCount >= P->cc_exile_time) transition(PS_CC_PU_SCAN);
                    else if (!cacheValid[lastCarriedType]) transition(PS_CC_PU_SCAN);
                    break;
                case PS_CC_SPIT_OUT:
                    if (stepCount - startCount >= P->cc_spit_out_time) transition(PS_CC_PU_SCAN);
                    break;
            }
        }
        if (state == PS_CC_PU_SCAN) {
            float randomTurn = P->cc_st_dev * (float)random->nextGaussian();
            movement = applyVFH(vfh, lm, randomTurn, ignorePucks());
        } else if (state == PS_CC_PU_TARGET) {
            movement = MovementCommand(1, (float)std::atan2((double)target.centroid.y, (double)target.centroid.x));
        } else if (state == PS_CC_HOMING) {
            float errorAngle = (float)Pose::getAngleFromTo(robotPose, cachePoints[carriedType]);
            lm.postFilterOccupancy(carriedType);
            movement = applyVFH(vfh, lm, errorAngle, ignorePucks());
        } else if (state == PS_CC_DE_PUSH) {
            movement = MovementCommand(1, 0);
        } else if (state == PS_CC_DE_BACKUP) {
            movement = MovementCommand(-1, 0);
        } else if (state == PS_CC_EXILE) {
            float errorAngle = (float)Pose::getAngleFromTo(robotPose, cachePoints[lastCarriedType]);
            errorAngle = (float)constrainAngle(errorAngle + M_PI);
            movement = applyVFH(vfh, lm, errorAngle, ignorePucks());
        } else if (state == PS_CC_SPIT_OUT) {
            float randomTurn = P->cc_st_dev * (float)random->nextGaussian();
            movement = MovementCommand(-1, randomTurn);
        }
        if (carrying) lastCarriedType = carriedType;
        if (stepCount % 100 == 0)   // :531-551 "Periodically save cache points to disk"
            for (int k = 0; k < PS_MAX_COLOURS; k++) {
                savedCache[k][0] = cacheValid[k] ? (float)cachePoints[k].x : std::numeric_limits<float>::quiet_NaN();
                savedCache[k][1] = cacheValid[k] ? (float)cachePoints[k].y : std::numeric_limits<float>::quiet_NaN();
            }
        updateStateCounts(stepCount);
    }
};

// src/controllers/ProbSeekController.java:176-308. Shares ExtremaSelector / ClusterTargetSelector / matchTarget with
// CacheCons (inherited); its property keys (ProbSeekController.K1, K2, PICK_UP_TIME, PUSH_IN_TIME, BACKUP_TIME,
// PREDICTION_DISTANCE_SQD, ST_DEV) have the CacheCons names and defaults and travel in the cc_* fields.
struct ProbSeekController : CacheConsController {
    int randomTurnDir = 0, randomTurnTime = 0;
    static const int TURNTIME_MIN = ONE_EIGHTY_TIME / 2;
    ProbSeekController(const ps_config* cfg, JavaRandom* rng) : CacheConsController(cfg, rng, false) {}
    void initiateRandomTurn(const LocalMap& lm) {  // :303-308
        randomTurnDir = lm.getFreeerSide();
        randomTurnTime = TURNTIME_MIN + random->nextInt(ONE_EIGHTY_TIME - TURNTIME_MIN + 1);
    }
    void computeDesired(LocalMap& lm, const Pose&, int stepCount_) override {  // :176-279
        bool carrying = lm.carrying;
        stepCount = stepCount_;
        switch (state) {
            case PS_PS_PU_SCAN:
                if (carrying) transition(PS_PS_DE_SCAN);
                else {
                    haveTarget = selectTarget(lm, carrying, !carrying, &target);
                    if (haveTarget) transition(PS_PS_PU_TARGET);
                }
                break;
            case PS_PS_PU_TARGET:
                if (carrying) transition(PS_PS_DE_SCAN);
                else {
                    Cluster nt;
                    haveTarget = matchTarget(lm, target, P->cc_prediction_distance_sqd, &nt);
                    if (haveTarget) target = nt;
                    if (!haveTarget) transition(PS_PS_PU_SCAN);
                    else if (stepCount - startCount > P->cc_pick_up_time) transition(PS_PS_PU_SCAN);
                }
                break;
            case PS_PS_DE_SCAN:
                if (!carrying) transition(PS_PS_PU_SCAN);
                else {
                    haveTarget = selectTarget(lm, carrying, !carrying, &target);
                    if (haveTarget) transition(PS_PS_DE_TARGET);
                }
                break;
            case PS_PS_DE_TARGET:
                if (!carrying) transition(PS_PS_PU_SCAN);
                else if (lm.carriedCluster >= 0 && lm.rawClusters[lm.carriedCluster].size > 1) transition(PS_PS_DE_PUSH);
                else {
                    Cluster nt;
                    haveTarget = matchTarget(lm, target, P->cc_prediction_distance_sqd, &nt);
                    if (haveTarget) target = nt;
                    if (!haveTarget) transition(PS_PS_DE_SCAN);
                    else if (stepCount - startCount > P->ps_placement_time) transition(PS_PS_DE_SCAN);
                }
                break;
            case PS_PS_DE_PUSH:
                if (stepCount - startCount >= P->cc_push_in_time) transition(PS_PS_DE_BACKUP);
                break;
            case PS_PS_DE_BACKUP:
                if (stepCount - startCount >= P->cc_backup_time) {
                    transition(PS_PS_DE_TURN);
                    initiateRandomTurn(lm);
                }
                break;
            case PS_PS_DE_TURN:
                if (stepCount - startCount >= randomTurnTime) transition(PS_PS_PU_SCAN);
                break;
        }
        if (state == PS_PS_PU_SCAN || state == PS_PS_DE_SCAN) {
            float randomTurn = P->cc_st_dev * (float)random->nextGaussian();
            movement = applyVFH(vfh, lm, randomTurn, false);
        } else if (state == PS_PS_PU_TARGET || state == PS_PS_DE_TARGET) {
            movement = MovementCommand(1, (float)std::atan2((double)target.centroid.y, (double)target.centroid.x));
        } else if (state == PS_PS_DE_PUSH) {
            movement = MovementCommand(1, 0);
        } else if (state == PS_PS_DE_BACKUP) {
            movement = MovementCommand(-1, 0);
        } else if (state == PS_PS_DE_TURN) {
            movement = MovementCommand(0, (float)(M_PI / 2 * randomTurnDir));
        }
        updateStateCounts(stepCount);
    }
};

struct BHDController : Controller {  // src/controllers/BHDController.java:126-191
    const ps_config* P;
    int stepCount = 0;
    int randomTurnDir = 0, randomTurnTime = 0;
    static const int TURNTIME_MIN = ONE_EIGHTY_TIME / 2;
    BHDController(const ps_config* cfg, JavaRandom* rng) : P(cfg) {
        random = rng;
        vfh.SAFETY_DISTANCE_GAMMA = cfg->vfh_safety_gamma;
        vfh.SAFETY_DISTANCE_MASK = cfg->vfh_safety_mask;
        vfh.TAU_LO = cfg->vfh_tau_lo * VFHPlus::A;
        vfh.TAU_HI = cfg->vfh_tau_hi * VFHPlus::A;
    }
    void transition(int s) {
        state = s;
        startCount = stepCount;
    }
    void initiateRandomTurn(const LocalMap& lm) {
        randomTurnDir = lm.getFreeerSide();
        randomTurnTime = TURNTIME_MIN + random->nextInt(ONE_EIGHTY_TIME - TURNTIME_MIN + 1);
    }
    void computeDesired(LocalMap& lm, const Pose&, int stepCount_) override {
        stepCount = stepCount_;
        switch (state) {
            case PS_BHD_FORWARD:
                if (lm.carriedCluster >= 0 && lm.rawClusters[lm.carriedCluster].size > 1) transition(PS_BHD_PUSH);
                else {
                    float result;
                    bool ok = vfh.computeTurnAngle(lm, 0, true, &result);
                    if (!ok || std::fabs(result) > M_PI / 4) {
                        transition(PS_BHD_TURN);
                        initiateRandomTurn(lm);
                    }
                }
                break;
            case PS_BHD_PUSH:
                if (stepCount - startCount >= P->bhd_push_in_time) transition(PS_BHD_BACKUP);
                break;
            case PS_BHD_BACKUP:
                if (stepCount - startCount >= P->bhd_backup_time) {
                    transition(PS_BHD_TURN);
                    initiateRandomTurn(lm);
                }
                break;
            case PS_BHD_TURN:
                if (stepCount - startCount >= randomTurnTime) transition(PS_BHD_FORWARD);
                break;
        }
        if (state == PS_BHD_FORWARD) movement = MovementCommand(1, 0);
        else if (state == PS_BHD_PUSH) movement = MovementCommand(1, 0);
        else if (state == PS_BHD_BACKUP) movement = MovementCommand(-1, 0);
        else if (state == PS_BHD_TURN) movement = MovementCommand(0, (float)(M_PI / 2 * randomTurnDir));
        updateStateCounts(stepCount);
    }
};

// ---- arena.Robot / arena.Puck / arena.DotPuck ----
struct Robot {
    jb::Body* body = nullptr;
    // PS_RNG_PER_ROBOT: the robot's own java.util.Random(seed * 1000003 + 7919 * (index + 1)) (an engine option for throughput
    // runs; the reference shares one Random per experiment, Experiment.java:22). On the heap: controllers keep a pointer to it.
    std::unique_ptr<JavaRandom> rng;
    std::unique_ptr<Controller> controller;
    LocalMap localMap;
    std::vector<uint8_t> image;  // STCameraImage.pixels[x][y] at x*H + y
};
struct Puck {
    jb::Body* body = nullptr;
    int colour = 0;
};

inline void createRobotFixtures(jb::World& w, jb::Body* body) {  // Robot.java:95-163
    jb::FixtureDef fd;
    fd.density = 5.0f;
    fd.categoryBits = 0x01;
    fd.maskBits = 0x07;
    float shiftX = -3.88f;
    float rightX = shiftX + 6.4f;
    Vec2 bodyVertices[5] = {Vec2(shiftX - 8.7f, 0.0f), Vec2(shiftX - 7.7f, -5.7f), Vec2(rightX, -5.7f), Vec2(rightX, 5.7f), Vec2(shiftX - 7.7f, 5.7f)};
    fd.shape = jb::Shape::polygon(bodyVertices, 5);
    fd.userKind = jb::U_ROBOT;
    w.createFixture(body, fd);
    fd.userKind = jb::U_HIDDEN;
    float innerShoulderY = 4.5f;
    Vec2 top[3] = {Vec2(rightX, innerShoulderY), Vec2(shiftX + 11.5f, 3.29f), Vec2(rightX, 5.7f)};
    Vec2 bot[3] = {Vec2(rightX, -5.7f), Vec2(shiftX + 11.5f, -3.29f), Vec2(rightX, -innerShoulderY)};
    fd.shape = jb::Shape::polygon(top, 3);
    w.createFixture(body, fd);
    fd.shape = jb::Shape::polygon(bot, 3);
    w.createFixture(body, fd);
    float startDegrees = 65.0f;
    float stopDegrees = 360.0f - startDegrees;
    createArc(w, body, shiftX + 10.0f, 0.0f, 3.6f, startDegrees * TO_RADf, stopDegrees * TO_RADf, 10, fd);
    fd.categoryBits = 0x02;
    fd.maskBits = 0x03;
    fd.shape = jb::Shape::circle(3.6f, Vec2(shiftX + 10.0f, 0.0f));
    w.createFixture(body, fd);
}

// ---- arena.Arena + arena.RunOffline.step ----
struct Arena {
    ps_config cfg;
    const Calibration* calib;
    JavaRandom rng;
    int64_t episodeSeed = 0;
    std::unique_ptr<jb::World> world;
    Enclosure enclosure;
    std::vector<Robot> robots;
    std::vector<Puck> pucks;
    int stepCount = 0, updateCount = 0;
    // the record Arena.step stores every STORAGE_INTERVAL think steps (Arena.java:223-243) and analysis.Analyze's per-repetition
    // statistics over those records (Analyze.java:176-218)
    int savedStep = -1;
    std::vector<float> savedPoses, savedPucks;
    ps_arena_analysis analysis = {0, -1, 0.0, 0.0, 0.0};
    float analysisThreshold = 1.5f, analysisTarget = 50.0f;

    Arena(const ps_config& c, const Calibration* cal) : cfg(c), calib(cal), analysisThreshold(c.cluster_distance_threshold) {}
    // PositionList.getClusterStats per colour (PositionList.java:85-96) -> Analyze's percent completion (Analyze.java:190-197)
    double percentCompletion(float threshold) {
        double sumMax = 0;
        for (int k = 0; k < cfg.n_puck_types; k++) {
            std::vector<Vec2> pos;
            for (const Puck& p : pucks)
                if (p.colour == k) pos.push_back(p.body->xf.position);
            std::vector<Cluster> cl;
            LocalMap::extractClusters(pos, 0, cl, threshold);
            int mx = 0;
            for (const Cluster& c : cl) mx = std::max(mx, c.size);
            sumMax += mx;
        }
        return cfg.n_pucks > 0 ? 100.0 * sumMax / (double)cfg.n_pucks : 0.0;
    }
    void storeRecord() {  // Arena.java:223-243, then one row of Analyze.computeEnsembleData
        savedStep = stepCount;
        savedPoses.clear();
        savedPucks.clear();
        for (Robot& r : robots) {
            savedPoses.push_back(r.body->xf.position.x);
            savedPoses.push_back(r.body->xf.position.y);
            savedPoses.push_back(r.body->sweep.a);
        }
        for (const Puck& p : pucks) {
            savedPucks.push_back(p.body->xf.position.x);
            savedPucks.push_back(p.body->xf.position.y);
        }
        double pc = percentCompletion(analysisThreshold);
        analysis.n_snapshots++;
        analysis.completion = pc;
        if (analysis.steps_to_completion == -1 && pc >= (double)analysisTarget) analysis.steps_to_completion = stepCount;
        analysis.twc_num += stepCount * pc / 100.0;
        analysis.twc_den += stepCount;
    }

    void createPuck(int colour, float minX, float maxX) {  // Arena.java:143-161
        Vec2 pos;
        float closest = 0;
        do {
            if (!enclosure.getRandomInside(PUCK_MIN_BOUNDARY_DISTANCE, &pos)) return;
            if (pos.x < minX || pos.x > maxX) continue;
            closest = INFINITY;
            for (const Puck& p : pucks)
                if (jb::distance(p.body->xf.position, pos) < closest) closest = jb::distance(p.body->xf.position, pos);
        } while (closest < PUCK_WIDTH);
        float randAngle = (float)(2.0f * M_PI * rng.nextFloat());
        placePuck(pos.x, pos.y, randAngle, colour);
    }
    void placePuck(float x, float y, float angle, int colour) {  // Puck.java:26-49, DotPuck.java:22-34
        jb::BodyDef bd;
        bd.type = jb::DYNAMIC_BODY;
        bd.position = Vec2(x, y);
        bd.angle = angle;
        bd.linearDamping = 10.0f;
        bd.angularDamping = 10.0f;
        jb::Body* body = world->createBody(bd);
        body->userKind = jb::U_PUCK;
        body->colour = colour;
        jb::FixtureDef fd;
        fd.density = 5.0f;
        fd.categoryBits = 0x04;
        fd.maskBits = 0x05;
        fd.shape = jb::Shape::circle(PUCK_WIDTH / 2);
        fd.userKind = jb::U_HIDDEN;
        world->createFixture(body, fd);
        fd.shape = jb::Shape::circle(INNER_WIDTH / 2);
        fd.userKind = jb::U_PUCK;
        fd.colour = colour;
        world->createFixture(body, fd);
        Puck p;
        p.body = body;
        p.colour = colour;
        pucks.push_back(p);
    }
    void createRobot(bool presetCaches) {  // Arena.java:163-181
        Vec2 pos;
        float closest = 0;
        int attempts = 0;
        do {
            if (!enclosure.getRandomInside(ROBOT_MIN_BOUNDARY_DISTANCE, &pos)) return;
            closest = INFINITY;
            for (const Robot& r : robots)
                if (jb::distance(r.body->xf.position, pos) < closest) closest = jb::distance(r.body->xf.position, pos);
            attempts++;
        } while (attempts < 1000 && closest < 2 * SRV_LENGTH);
        float randAngle = (float)(2.0f * M_PI * rng.nextFloat());
        placeRobot(pos.x, pos.y, randAngle, presetCaches);
    }
    void placeRobot(float x, float y, float angle, bool presetCaches) {  // Robot.java:65-93
        jb::BodyDef bd;
        bd.type = jb::DYNAMIC_BODY;
        bd.position = Vec2(x, y);
        bd.angle = angle;
        bd.linearDamping = 10.0f;
        bd.angularDamping = 10.0f;
        jb::Body* body = world->createBody(bd);
        body->userKind = jb::U_ROBOT;
        createRobotFixtures(*world, body);
        robots.emplace_back();
        Robot& r = robots.back();
        r.body = body;
        r.localMap.init(calib);
        r.localMap.CLUSTER_DISTANCE_THRESHOLD = cfg.cluster_distance_threshold;
        r.localMap.CARRY_THRESHOLD_LO = cfg.carry_threshold_lo;
        r.localMap.CARRY_THRESHOLD_HI = cfg.carry_threshold_hi;
        r.image.assign((size_t)calib->imageWidth * calib->imageHeight, PS_T_NOTHING);
        r.rng.reset(new JavaRandom(episodeSeed * 1000003LL + 7919LL * (int64_t)robots.size()));
        JavaRandom* stream = cfg.rng_mode == PS_RNG_PER_ROBOT ? r.rng.get() : &rng;
        if (cfg.controller == PS_CTRL_BHD) r.controller.reset(new BHDController(&cfg, stream));
        else if (cfg.controller == PS_CTRL_PROBSEEK) r.controller.reset(new ProbSeekController(&cfg, stream));
        else r.controller.reset(new CacheConsController(&cfg, stream, presetCaches));
        // VFHPlus re-runs initDataStructures on its first computeTurnAngle (propertiesUpdated is set by the
        // constructor, VFHPlus.java:167,227-230); nothing touches the histograms before that, so doing it here
        // is equivalent and lets set_state restore binaryHistogram at any time.
        r.controller->vfh.initDataStructures(r.localMap);
        r.controller->vfh.inited = true;
    }
    void newWorld() {
        jb::Switches sw;
        sw.sincosLut = cfg.sincos_lut != 0;
        sw.toiEnabled = cfg.toi_enabled != 0;
        world.reset(new jb::World());
        world->setSwitches(sw);
        robots.clear();
        pucks.clear();
        stepCount = 0;
        updateCount = 0;
        savedStep = -1;
        savedPoses.clear();
        savedPucks.clear();
        analysis = ps_arena_analysis{0, -1, 0.0, 0.0, 0.0};
    }
    // RunOffline.recreateArena (:97-113) + new Arena(world, null, false) (Arena.java:45-120) for Experiment(seed)
    void reset(int64_t seed) {
        episodeSeed = seed;
        rng.setSeed(seed);
        newWorld();
        robots.reserve(cfg.n_robots);
        enclosure.build(*world, cfg.enclosure_scale, &rng);
        int nPucksPerType = cfg.n_pucks / cfg.n_puck_types;
        float w = enclosure.WIDTH;
        for (int k = 0; k < cfg.n_puck_types; k++)
            for (int i = 0; i < nPucksPerType; i++) createPuck(k, -w / 2, w / 2);
        for (int i = 0; i < cfg.n_robots; i++) createRobot(i < cfg.n_informed_robots);
        // Canonical transform: JBox2D keeps the creation transform (position = BodyDef.position) until the first
        // synchronizeTransform, which recomputes position = c - R * localCenter and may move a robot's origin by one
        // ulp. The state exchange format carries (c, a) only, so the oracle adopts the recomputed transform right
        // after construction (fixture AABBs and the first pair search above still use the creation transform).
        for (jb::Body* b = world->bodyList; b; b = b->next)
            if (b->type == jb::DYNAMIC_BODY) b->synchronizeTransform();
    }
    // sensors.PointSensor.pointSense (src/sensors/PointSensor.java:85-114)
    uint8_t pointSense(const jb::Body* robotBody, const Vec2& posWrtBody) const {
        Vec2 globalPos = jb::mul(robotBody->xf, posWrtBody);
        if (!enclosure.inFreeSpace(globalPos, 0)) return PS_T_WALL;
        for (const jb::Body* b = world->bodyList; b; b = b->next) {
            if (b == robotBody || jb::distanceSquared(b->xf.position, globalPos) > RADIUS_OF_LARGEST_OBJECT_SQD) continue;
            for (const jb::Fixture* f = b->fixtureList; f; f = f->next) {
                if (f->userKind != jb::U_HIDDEN && globalPos.x <= f->aabb.upper.x && globalPos.x >= f->aabb.lower.x && globalPos.y <= f->aabb.upper.y &&
                    globalPos.y >= f->aabb.lower.y && f->shape.testPoint(b->xf, globalPos)) {
                    if (f->userKind == jb::U_WALL) return PS_T_WALL;
                    if (f->userKind == jb::U_PUCK) return (uint8_t)f->colour;
                    return PS_T_ROBOT;
                }
            }
        }
        return PS_T_NOTHING;
    }
    // sensors.GridCamera.sense (src/sensors/GridCamera.java:48-83)
    void cameraSense(Robot& r) const {
        int W = calib->imageWidth, H = calib->imageHeight;
        for (int yp = 0; yp < H; yp++)
            for (int xp = 0; xp < W; xp++) {
                const CalibPixel* p = calib->get(xp, yp);
                if (!p || p->gripperBody) r.image[(size_t)xp * H + yp] = PS_T_HIDDEN;
                else r.image[(size_t)xp * H + yp] = pointSense(r.body, Vec2(p->Xr, p->Yr));
            }
    }
    Pose getPose(const Robot& r) const {  // SimAPS.java:14-18
        Pose p;
        p.x = r.body->xf.position.x;
        p.y = r.body->xf.position.y;
        p.theta = r.body->sweep.a;
        return p;
    }
    void move(Robot& r) {  // Robot.java:175-182
        Vec2 f = r.body->getWorldVector(Vec2(r.controller->movement.forwards, 0.0f));
        Vec2 p = r.body->getWorldPoint(r.body->sweep.localCenter);
        r.body->applyForce(f, p);
        r.body->applyTorque(r.controller->movement.torque);
    }
    void senseAll() {
        for (Robot& r : robots) {
            cameraSense(r);
            r.localMap.update(r.image);
        }
    }
    // Arena.getRandomPuckPos (Arena.java:318-329): a free-space point at least one puck width from every robot origin
    bool getRandomPuckPos(Vec2* out) {
        float closest = 0;
        Vec2 pos;
        do {
            if (!enclosure.getRandomInside(PUCK_MIN_BOUNDARY_DISTANCE, &pos)) return false;   // null: the arena is blocked full
            closest = INFINITY;
            for (const Robot& r : robots)
                if (jb::distance(r.body->xf.position, pos) < closest) closest = jb::distance(r.body->xf.position, pos);
        } while (closest < PUCK_WIDTH);
        *out = pos;
        return true;
    }
    void teleportPuckRandomly(Puck& p) {
        Vec2 pos;
        if (getRandomPuckPos(&pos)) world->setTransform(p.body, pos, 0);
    }
    // Arena.shiftPucksThroughCentre + shiftPucks (Arena.java:334-343, 354-373)
    void shiftPucksThroughCentre() {
        Vec2 centroid;
        for (Puck& p : pucks) centroid += p.body->xf.position;
        centroid.normalize();
        centroid *= -100;
        for (Puck& p : pucks) {
            Vec2 pos = p.body->xf.position + centroid;
            if (!enclosure.inFreeSpace(pos, PUCK_MIN_BOUNDARY_DISTANCE)) {
                teleportPuckRandomly(p);
            } else {
                float closest = INFINITY;
                for (const Robot& r : robots)
                    if (jb::distance(r.body->xf.position, pos) < closest) closest = jb::distance(r.body->xf.position, pos);
                if (closest < PUCK_WIDTH) teleportPuckRandomly(p);
                else world->setTransform(p.body, pos, 0);
            }
        }
    }
    // Arena.shufflePucks (Arena.java:345-348)
    void shufflePucks() {
        for (Puck& p : pucks) teleportPuckRandomly(p);
    }
    // the head of Arena.step (Arena.java:185-192)
    void perturb() {
        if (cfg.puck_shift_step != -1 && stepCount % cfg.puck_shift_step == 0) shiftPucksThroughCentre();
        if (cfg.puck_shuffle_step != -1 && stepCount % cfg.puck_shuffle_step == 0) shufflePucks();
    }
    void arenaStep() {  // Arena.java:183-267 with (true,true,true,false,0,false)
        perturb();
        for (Robot& r : robots) {
            cameraSense(r);
            r.localMap.update(r.image);
            if (cfg.controller != PS_CTRL_EXTERNAL) r.controller->computeDesired(r.localMap, getPose(r), stepCount);
            move(r);
        }
        if (stepCount % 100 == 0) storeRecord();   // Arena.STORAGE_INTERVAL
        stepCount++;
    }
    void coast() {  // Arena.java:310-316
        for (Robot& r : robots) move(r);
    }
    void tick() {  // RunOffline.step (:72-95) minus the experiment bookkeeping
        if (updateCount % cfg.step_interval == 0) arenaStep();
        else coast();
        updateCount++;
        world->step(cfg.dt, cfg.vel_iters, cfg.pos_iters);
    }
};

}  // namespace oracle
