/*
 * puckswarm_b200.h -- C-ABI of the B200 batched engine for PuckSwarm's per-tick hot path.
 *
 * The reference (BOTSlab/PuckSwarm, pure Java) has no FFI; the seams this ABI replaces are
 *   - arena.RunOffline.step()                 src/arena/RunOffline.java:72-95   -> ps_step
 *   - new World(...) + new Arena(world,..)    src/arena/RunOffline.java:97-113,
 *                                             src/arena/Arena.java:45-120       -> ps_create + ps_reset
 *   - Arena.step / Arena.coast                src/arena/Arena.java:183-267,310-316 (inside ps_step)
 *   - World.step(1/14f, 3, 100)               src/arena/RunOffline.java:94      (inside ps_step)
 *   - sensors.Suite.sense/getCameraImage/getLocalMap/getAPS
 *                                             src/sensors/Suite.java:5-43       -> ps_get_observations
 *   - controllers.Controller.computeDesired/getForwards/getTorque
 *                                             src/controllers/Controller.java:9-31
 *                                                                               -> on device (CacheCons, BHD) or
 *                                                                                  ps_step_external (host controllers)
 *   - GridCalibration(dir, 160, 120)          src/sensors/GridCalibration.java:53-114 -> ps_load_calibration
 *   - Experiment.getProperty(key, default)    src/experiment/Experiment.java:76-90    -> ps_config fields
 *   - PositionList.getClusterStats / Analyze  src/arena/PositionList.java:85-96,
 *                                             src/analysis/Analyze.java:187-218 -> ps_compute_stats
 *
 * Conventions: plain C, opaque handle, caller-allocated arrays, SoA and arena-major. Every call returns
 * 0 (PS_OK) or a negative PS_E* code; ps_last_error() gives the message of the last failure on this
 * thread. A handle is confined to one host thread and owns one CUDA device + stream. Nothing here
 * falls back to the CPU: without a CUDA device ps_create fails with PS_E_NODEVICE.
 *
 * The same structs are the state-exchange format of the CPU oracle (oracle/), which is test
 * infrastructure only.
 */
#ifndef PUCKSWARM_B200_H
#define PUCKSWARM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PS_ABI_VERSION 3

/* ---- fixed sizes of the scene ---- */
#define PS_MAX_COLOURS      8    /* SensedType.NPUCK_COLOURS, src/sensors/SensedType.java:39 */
#define PS_ROBOT_FIXTURES   14   /* src/arena/Robot.java:95-163 */
#define PS_PUCK_FIXTURES    2    /* src/arena/DotPuck.java:22-34 */
#define PS_WALL_FIXTURES    84   /* src/utils/FixtureUtils.java:51-74 */
#define PS_MAX_LM_PUCKS     96   /* capacity of one robot's perceived-puck list (all colours) */
#define PS_N_SECTORS        73   /* VFHPlus.N, src/localmap/VFHPlus.java:81 */

/* ---- SensedType ordinals, src/sensors/SensedType.java:9-23 ---- */
enum { PS_T_NOTHING = 8, PS_T_WALL = 9, PS_T_ROBOT = 10, PS_T_HIDDEN = 11 };

/* ---- error codes ---- */
enum {
    PS_OK = 0,
    PS_E_INVALID   = -1,  /* bad argument / config */
    PS_E_NODEVICE  = -2,  /* no CUDA device or wrong architecture: there is no CPU fallback */
    PS_E_CUDA      = -3,  /* a CUDA runtime call failed */
    PS_E_NOCALIB   = -4,  /* ps_load_calibration has not been called */
    PS_E_CAPACITY  = -5,  /* a per-arena capacity (contacts, perceived pucks) overflowed on device */
    PS_E_NOMEM     = -6
};

/* controllerType property, src/controllers/ControllerUtils.java:6-41 */
enum { PS_CTRL_CACHECONS = 0, PS_CTRL_BHD = 1, PS_CTRL_EXTERNAL = 2, PS_CTRL_PROBSEEK = 3 };
/* shared java.util.Random per experiment (src/experiment/Experiment.java:22) or one stream per robot */
enum { PS_RNG_JAVA_SHARED = 0, PS_RNG_PER_ROBOT = 1 };
/* CacheConsController.CacheSelectionOption / CacheSeparationOption, CacheConsController.java:136-177 */
enum { PS_SEL_PROB_ABSOLUTE = 0, PS_SEL_PROB_RELATIVE = 1, PS_SEL_RELATIVE = 2, PS_SEL_RELATIVE_FORGET = 3 };
enum { PS_SEP_IGNORE = 0, PS_SEP_NULLIFY_OLD = 1, PS_SEP_ACCEPT_ISOLATED = 2, PS_SEP_ACCEPT_LARGER = 3 };
/* CacheConsController.State / BHDController.State ordinals */
enum { PS_CC_PU_SCAN = 0, PS_CC_PU_TARGET, PS_CC_HOMING, PS_CC_DE_PUSH, PS_CC_DE_BACKUP, PS_CC_EXILE, PS_CC_SPIT_OUT };
enum { PS_BHD_FORWARD = 0, PS_BHD_PUSH, PS_BHD_BACKUP, PS_BHD_TURN };
/* ProbSeekController.State ordinals, src/controllers/ProbSeekController.java:27-34 */
enum { PS_PS_PU_SCAN = 0, PS_PS_PU_TARGET, PS_PS_DE_SCAN, PS_PS_DE_TARGET, PS_PS_DE_PUSH, PS_PS_DE_BACKUP, PS_PS_DE_TURN };

/*
 * Configuration: one POD filled from the .properties keys. ps_config_defaults()
 * sets every field to the reference default.
 */
typedef struct ps_config {
    int32_t abi_version;          /* PS_ABI_VERSION */
    int32_t n_arenas;
    int32_t n_robots;             /* Arena.nRobots (4)            src/arena/Arena.java:98-99 */
    int32_t n_pucks;              /* Arena.nPucks (10)            src/arena/Arena.java:58-59 */
    int32_t n_puck_types;         /* Arena.nPuckTypes (2)         src/arena/Arena.java:60-61 */
    int32_t n_informed_robots;    /* Arena.nInformedRobots (0)    src/arena/Arena.java:101-102 */
    int32_t controller;           /* controllerType ("CacheCons") */
    int32_t rng_mode;             /* PS_RNG_JAVA_SHARED */
    int32_t step_interval;        /* PuckSwarmTest.STEP_INTERVAL (3) src/arena/PuckSwarmTest.java:30 */
    int32_t vel_iters;            /* 3    src/arena/RunOffline.java:94 */
    int32_t pos_iters;            /* 100  src/arena/RunOffline.java:94 */
    float   dt;                   /* 1/14f */
    float   enclosure_scale;      /* RoundedRectangleEnclosure.scale (1) */
    int32_t sincos_lut;           /* 1: JBox2D Settings.SINCOS_LUT_ENABLED semantics */
    int32_t toi_enabled;          /* 1: World.solveTOI against the static wall body */
    int32_t max_contacts;         /* per-arena contact-cache capacity; 0 = derive from body counts */
    float   cluster_distance_threshold;   /* LocalMap.CLUSTER_DISTANCE_THRESHOLD (1.5) LocalMap.java:74 */
    float   carry_threshold_lo;   /* LocalMap.CARRY_THRESHOLD_LO (0.1) */
    float   carry_threshold_hi;   /* LocalMap.CARRY_THRESHOLD_HI (0.4) */
    float   vfh_safety_gamma;     /* VFHPlus.SAFETY_DISTANCE_GAMMA (2) */
    float   vfh_safety_mask;      /* VFHPlus.SAFETY_DISTANCE_MASK (5) */
    float   vfh_tau_lo;           /* VFHPlus.TAU_LO (0.7) */
    float   vfh_tau_hi;           /* VFHPlus.TAU_HI (0.85) */
    /* CacheConsController.* (src/controllers/CacheConsController.java:250-289) */
    float   cc_k1;                /* K1 (1) */
    float   cc_k2;                /* K2 (8) */
    int32_t cc_pick_up_time;      /* 20 */
    int32_t cc_push_in_time;      /* 1 */
    int32_t cc_backup_time;       /* 5 */
    float   cc_prediction_distance_sqd;  /* 100 */
    float   cc_st_dev;            /* ST_DEV (0.5) */
    int32_t cc_exile_time;        /* 20 */
    int32_t cc_spit_out_time;     /* 10 */
    int32_t cc_cache_selection;   /* PS_SEL_RELATIVE */
    float   cc_k_absolute;        /* 40 */
    float   cc_k_relative;        /* 8 */
    float   cc_forget_prob;       /* 0.001 */
    int32_t cc_cache_separation;  /* PS_SEP_ACCEPT_LARGER */
    float   cc_home_separation_distance; /* 50 */
    int32_t cc_ignore_pucks;      /* 0 */
    /* BHDController.* (src/controllers/BHDController.java:76-79) */
    int32_t bhd_push_in_time;     /* 1 */
    int32_t bhd_backup_time;      /* 5 */
    /* ProbSeekController.* (src/controllers/ProbSeekController.java:98-121): K1, K2, PICK_UP_TIME, PUSH_IN_TIME, BACKUP_TIME,
     * PREDICTION_DISTANCE_SQD and ST_DEV have the names and defaults of the CacheCons keys and travel in the cc_* fields above
     * when controller == PS_CTRL_PROBSEEK; the one key of its own: */
    int32_t ps_placement_time;    /* PLACEMENT_TIME (40) */
    /* Perturbations at the start of Arena.step, every think step whose stepCount is a multiple of the value; -1 = never
     * (src/arena/Arena.java:186-192, 318-373: shiftPucksThroughCentre / shufflePucks through Body.setTransform) */
    int32_t puck_shift_step;      /* Arena.puckShiftStep (-1) */
    int32_t puck_shuffle_step;    /* Arena.puckShuffleStep (-1) */
} ps_config;

/*
 * Persistent per-robot sensor + controller state (everything that survives from one think-step to
 * the next; the local map itself is rebuilt every sense, LocalMap.java:183-208).
 */
typedef struct ps_robot_state {
    double   cache_x[PS_MAX_COLOURS];     /* CacheConsController.cachePoints (Pose x), :79 */
    double   cache_y[PS_MAX_COLOURS];
    double   rng_next_gaussian;           /* per-robot stream (PS_RNG_PER_ROBOT only) */
    uint64_t rng_seed;
    float    remembered[PS_MAX_COLOURS];  /* rememberedCacheSizes, :84 */
    uint8_t  cache_valid[PS_MAX_COLOURS]; /* cachePoints[k] != null */
    int32_t  rng_have_gaussian;
    int32_t  ctrl_state;                  /* State ordinal */
    int32_t  start_count;
    int32_t  has_target;                  /* target != null (always a one-puck cluster) */
    int32_t  target_type;
    float    target_x, target_y;
    int32_t  last_carried_type;           /* -1 = null */
    int32_t  lm_carrying;                 /* LocalMap.carrying of the previous update (hysteresis, :238) */
    uint32_t vfh_binary[3];               /* VFHPlus.binaryHistogram as 73 bits */
    int32_t  vfh_last_sector;             /* VFHPlus.lastTurnSector */
    float    vfh_last_angle;              /* VFHPlus.lastTurnAngle */
    int32_t  bhd_turn_dir, bhd_turn_time; /* BHDController.randomTurnDir/Time */
    int32_t  state_counts[8];             /* recentStateCounts since the last storage interval */
    float    cmd_forwards, cmd_torque;    /* Controller.getForwards/getTorque, re-applied by coast */
    int32_t  pad_;
} ps_robot_state;

typedef struct ps_arena_state {
    double   rng_next_gaussian;   /* shared java.util.Random (PS_RNG_JAVA_SHARED) */
    uint64_t rng_seed;
    int32_t  rng_have_gaussian;
    int32_t  update_count;        /* RunOffline.updateCount (physics ticks done) */
    int32_t  step_count;          /* Arena.stepCount (think steps done) */
    int32_t  n_contacts;          /* live entries of the contact cache */
    float    inv_dt0;             /* World.m_inv_dt0 */
    int32_t  error_flags;         /* capacity overflows (the arena has left the reference's trajectory): bit0 contact cache (max_contacts),
                                     bit1 perceived pucks (PS_MAX_LM_PUCKS; reported by ps_sync / ps_get_error_flags from the local maps),
                                     bit2 touching constraints (max_contacts / 4), bit3 broad-phase pair buffer, bit4 wall-contact list */
} ps_arena_state;

/*
 * Full simulation state, for lock-step parity and checkpointing. All pointers are HOST arrays owned by
 * the caller; a NULL pointer skips that component. Indexing:
 *   NB = n_pucks + n_robots dynamic bodies, pucks first (creation order, src/arena/Arena.java:91-106)
 *   NP = 2*n_pucks + 14*n_robots dynamic proxies; global proxy id = 84 + dynamic index (walls are 0..83),
 *        i.e. fixture creation order, which is the broad-phase proxy order
 *   contact cache: entries in creation order; JBox2D's world contact list is this array reversed.
 */
typedef struct ps_state_view {
    float*          body;            /* [n_arenas][6][NB]: sweep.c.x, sweep.c.y, sweep.a, v.x, v.y, omega */
    float*          sense_aabb;      /* [n_arenas][4][NB]: Fixture.m_aabb (swept tight AABB of the last synchronizeFixtures) of the
                                        one fixture per body that PointSensor.possibleOverlap can reach: robot hull / puck inner disc
                                        (src/sensors/PointSensor.java:95-108,111-114) */
    float*          fat_aabb;        /* [n_arenas][4][NP]: broad-phase fat AABB lower.x, lower.y, upper.x, upper.y */
    uint32_t*       contact_key;     /* [n_arenas][max_contacts]: (proxyA << 16) | proxyB as stored in the contact */
    uint32_t*       contact_info;    /* [n_arenas][max_contacts]: bits0-1 manifold.pointCount, bits8-15 id0, bits16-23 id1 */
    float*          contact_impulse; /* [n_arenas][max_contacts][4]: normal0, tangent0, normal1, tangent1 */
    ps_robot_state* robot;           /* [n_arenas][n_robots] */
    ps_arena_state* arena;           /* [n_arenas] */
} ps_state_view;

/* One robot's LocalMap after update() (src/localmap/LocalMap.java:183-208). */
typedef struct ps_localmap {
    int32_t carrying;                 /* LocalMap.isCarrying() */
    int32_t carried_type;             /* SensedType ordinal, PS_T_NOTHING if none */
    float   carried_x, carried_y;     /* carriedV */
    int32_t carried_cluster;          /* index into clusters, -1 = null */
    int32_t n_pucks;                  /* pucks[0..7] concatenated, blob (seed-scan) order within a colour */
    int32_t n_clusters;               /* rawClusters, order of LocalMap.update */
    int32_t left_sum, right_sum;      /* getFreeerSide() tallies, :565-587 */
    int32_t overflow;                 /* perceived pucks exceeded PS_MAX_LM_PUCKS */
    float   puck_x[PS_MAX_LM_PUCKS], puck_y[PS_MAX_LM_PUCKS];
    float   cluster_x[PS_MAX_LM_PUCKS], cluster_y[PS_MAX_LM_PUCKS];   /* centroids */
    uint8_t puck_type[PS_MAX_LM_PUCKS];
    uint8_t puck_cluster[PS_MAX_LM_PUCKS];
    uint8_t puck_degree[PS_MAX_LM_PUCKS];      /* subgraph.degreeOf */
    uint8_t cluster_type[PS_MAX_LM_PUCKS];
    uint8_t cluster_size[PS_MAX_LM_PUCKS];
    uint8_t cluster_kept[PS_MAX_LM_PUCKS];     /* 1 = survived filterClusters (:375-437) */
} ps_localmap;

/* Observations of the latest sense (Suite.getCameraImage/getLocalMap/getAPS). NULL pointers are skipped. */
typedef struct ps_obs_view {
    uint8_t*     camera;     /* [n_arenas][n_robots][img_h*img_w], pixel (x,y) at y*img_w + x; SensedType ordinals */
    uint8_t*     occupancy;  /* [n_arenas][n_robots][occ_w*occ_h], cell (i,j) at i*occ_h + j */
    ps_localmap* localmap;   /* [n_arenas][n_robots] */
    float*       pose;       /* [n_arenas][n_robots][3]: SimAPS pose x, y, theta (SimAPS.java:14-18) */
} ps_obs_view;

/* Experiment statistics reduced over the arenas of the handle (PositionList.getClusterStats, src/arena/PositionList.java:85-96;
 * Analyze.java:176-259). The fields are grouped by the reduction that combines ranks, in this order, so that a multi-GPU reduce
 * is one collective per group: int64 sums | double sums | int64 minima | int64 maxima | double minima | double maxima. */
typedef struct ps_stats {
    /* --- int64, summed */
    int64_t n_arenas;
    int64_t cluster_size_hist[PS_MAX_COLOURS][64]; /* count of clusters by size (sizes >= 63 in the last bin) */
    int64_t ctrl_state_hist[8];                    /* robots per controller state */
    int64_t n_carrying;                            /* robots with LocalMap.carrying */
    int64_t n_caches;                              /* CacheCons: valid cache points over all robots */
    int64_t n_touching_points;                     /* live touching manifold points (C-hat of SURVEY 8d) */
    int64_t n_contacts;                            /* live cached contacts */
    int64_t n_completed;                           /* arenas whose stepsToCompletion != -1 (Analyze.java:247-256) */
    int64_t sum_steps_to_completion;               /* over the completed arenas */
    int64_t n_snapshots;                           /* storage-step records analysed, over all arenas */
    /* --- double, summed */
    double  sum_completion;                        /* sum over arenas of the CURRENT 100 * sum_k max cluster size_k / Arena.nPucks */
    double  sum_twc;                               /* sum over arenas of timeWeightedCompletion (Analyze.java:206-216) */
    /* --- int64 minimum / maximum over the completed arenas (INT64_MAX / -1 when none completed) */
    int64_t min_steps_to_completion;
    int64_t max_steps_to_completion;
    /* --- double minima, then maxima, over the arenas */
    double  min_completion, min_twc;
    double  max_completion, max_twc;
} ps_stats;

/* The reference's periodic record, kept for the latest storage step (Arena.stepCount % 100 == 0) of every arena: what Arena.step
 * writes to step%07d_robots.txt / _<COLOUR>_pucks.txt (Arena.java:223-243; poses after the perturbations, before World.step) and
 * what the on-device controllers would write to _R<i>_cachePoints.txt / _R<i>_stateCounts.txt (CacheConsController.java:313-333,
 * 531-551, BHDController.java:103-123). All pointers are HOST arrays of the caller; NULL skips a component. */
typedef struct ps_snapshot_view {
    int32_t* step_count;     /* [n_arenas]: Arena.stepCount of the record, -1 = none yet */
    float*   robot_pose;     /* [n_arenas][n_robots][3]: body.m_xf.position.x, .y, body.getAngle() */
    float*   puck_xy;        /* [n_arenas][P][2], pucks in creation order (colour k = pucks k*P/n_puck_types ..) */
    float*   cache_xy;       /* [n_arenas][n_robots][PS_MAX_COLOURS][2]: (float)cachePoints[k], NaN = null (CacheCons only) */
    int32_t* state_counts;   /* [n_arenas][n_robots][8]: recentStateCounts as saved, before they are cleared */
} ps_snapshot_view;

/* analysis.Analyze.computeEnsembleData for one repetition (Analyze.java:176-218), accumulated on the device at every storage
 * step: percent completion = 100 * sum_k max cluster size_k / Arena.nPucks over the recorded puck positions. */
typedef struct ps_arena_analysis {
    int32_t n_snapshots;          /* rows so far (storage steps 0, 100, 200, ...) */
    int32_t steps_to_completion;  /* first stepCount with completion >= target (TARGET_PERCENT_COMPLETION = 50), -1 = not reached */
    double  twc_num;              /* sum over rows of stepCount * completion / 100 */
    double  twc_den;              /* sum over rows of stepCount; timeWeightedCompletion = twc_num / twc_den */
    double  completion;           /* percent completion of the latest row */
} ps_arena_analysis;

typedef struct ps_engine* ps_handle;

const char* ps_last_error(void);
int  ps_abi_version(void);
/* sizeof() of the ABI structs as the library was compiled: 0 ps_config, 1 ps_robot_state, 2 ps_arena_state,
 * 3 ps_localmap, 4 ps_stats, 5 ps_state_view, 6 ps_obs_view (lets a foreign-language binding verify its layout) */
int  ps_struct_size(int which);
void ps_config_defaults(ps_config* cfg);

int ps_create(const ps_config* cfg, int device, ps_handle* out);
int ps_destroy(ps_handle h);

/* Calibration rows: xy[n][2] = pixel (x,y); XrYr[n][2] = ground-plane point in the robot frame;
 * flags bit0 = gripperBody, bit1 = gripperHole (GridCalibration.java:66-107). */
int ps_load_calibration(ps_handle h, const float* XrYr, const uint16_t* xy, const uint8_t* flags,
                        int n, int img_w, int img_h);

/* Spawn every arena from its seed exactly as new Experiment(seed) + new Arena(...) would (a24). */
int ps_reset(ps_handle h, const int64_t* seeds);

/* Partial views (NULL pointers) are allowed within these rules, enforced with PS_E_INVALID:
 *   - contact_key, contact_info, contact_impulse and arena (for n_contacts) travel together; n_contacts <= max_contacts and every
 *     key must name two different proxies of the arena;
 *   - update_count must be the same in every arena (the batch steps in lock step; the host owns the think schedule);
 *   - the first upload into a handle that was never reset must carry every component.
 * Not checkable, so the caller's duty: the cache must equal the set of filtered proxy pairs whose fat AABBs overlap (JBox2D's
 * invariant between steps; the solver's "new pair iff the boxes overlap now and did not before" relies on it). Uploading `body`
 * alone is therefore only valid for poses that still lie inside the uploaded / resident fat AABBs (e.g. the state just downloaded,
 * or velocity edits); to teleport a body upload body + sense_aabb + fat_aabb + the contact group of a consistent state. */
int ps_set_state(ps_handle h, const ps_state_view* v);
int ps_get_state(ps_handle h, ps_state_view* v);

/* n_ticks iterations of RunOffline.step(): {Arena.step | Arena.coast} then World.step. Asynchronous on the
 * handle's stream; ps_sync waits and reports device-side capacity errors. */
int ps_step(ps_handle h, int n_ticks);
/* Waits for the handle's work. Returns PS_E_CAPACITY (after waiting; the state stays readable) when any arena raised a capacity
 * flag: such an arena dropped a contact / constraint / pair / perceived puck and no longer follows the reference. */
int ps_sync(ps_handle h);
/* OR of all arenas' error_flags (bit1 also from the robots' last local maps) and the number of flagged arenas. */
int ps_get_error_flags(ps_handle h, int32_t* flags, int32_t* n_flagged);
/* Order the handle's main stream (ps_stream) after everything ps_step has queued so far, without waiting on the host:
 * an event recorded on that stream afterwards completes when the steps have. */
int ps_join(ps_handle h);

/* Host-controller path: run sense (camera + local map) for every robot now. */
int ps_sense(ps_handle h);
/* ... then one think interval (step_interval ticks) with host-supplied commands [n_arenas][n_robots]. */
int ps_step_external(ps_handle h, const float* forwards, const float* torque);

int ps_get_observations(ps_handle h, ps_obs_view* v);

/* Per-arena statistics reduced on device into one ps_stats (this process' arenas only). The device copy
 * (ps_stats_device_ptr, sizeof(ps_stats) bytes of int64/double) is what the host all-reduces over NCCL. */
int ps_compute_stats(ps_handle h, float cluster_threshold, ps_stats* out);
int ps_stats_device_ptr(ps_handle h, void** dptr, int64_t* n_bytes);
/* The same, reduced over the ranks of an NCCL communicator (ps_get_stats(h, out, allreduce)): nccl_comm is the
 * caller's ncclComm_t for this handle's device (NULL = this process only). One ncclAllReduce per reduction group of ps_stats
 * (sum / min / max over int64 and double) on the handle's stream, straight on the device copy; libnccl.so.2 is loaded on first use. */
int ps_get_stats(ps_handle h, float cluster_threshold, ps_stats* out, void* nccl_comm);

/* Latest storage-step record of every arena / the per-arena Analyze accumulators (see the structs). The clustering threshold and
 * the completion target default to ps_config.cluster_distance_threshold and 50 (Analyze.java:32); ps_analysis_config overrides them
 * and clears the accumulators. */
int ps_get_snapshot(ps_handle h, ps_snapshot_view* v);
int ps_get_analysis(ps_handle h, ps_arena_analysis* out /* [n_arenas] */);
int ps_analysis_config(ps_handle h, float cluster_threshold, float target_percent);

/* Engine introspection */
int ps_get_dims(ps_handle h, int32_t* n_bodies, int32_t* n_proxies, int32_t* max_contacts,
                int32_t* occ_w, int32_t* occ_h);
/* Diagnostics of the last World.step of every arena, [n_arenas][8] int32: 0 = max position iterations of an island,
 * 1 = TOI events, 2 = islands with contacts. */
int ps_get_step_stats(ps_handle h, int32_t* out);
/* Diagnostics of the last sense of every robot, [n_arenas][n_robots][8] int32: cycles / 16 of 0 = candidate cull, 1 = camera,
 * 2 = occupancy + blob labelling, 3 = occupancy (cycle counts are zero unless the library is built with PS_PHASE_CLOCKS);
 * 4 = candidate bodies examined summed over the robot's pixels, 5 = puck-coloured pixels, 6 = candidate bodies, 7 = blobs. */
int ps_get_sense_stats(ps_handle h, int32_t* out);
/* Diagnostics: islands with >= 3 contacts solved since the last reset, as an 8 x 8 histogram [position-iteration bucket][contact
 * bucket]; iterations <=1, <=3, <=6, <=12, <=25, <=50, <100, 100; contacts <=3, <=5, <=8, <=12, <=16, <=32, <=64, >64. */
int ps_get_island_hist(ps_handle h, int64_t* out /*[64]*/, int reset);
/* Diagnostics: census of stage B since the last reset. [0..9] islands per queue bucket (0 warp, 1-4 lane buckets, 5 big, 6 large,
 * 7 XL, 8 XXL, 9 wide lane slot), [10] ticks (of arena group 0's first arena); [16..31] warp-slot islands by duration in bins of
 * 2^18 cycles, [32..47] the same for the big islands; [48],[49] kilo-cycles and count of the warp-slot islands, [50],[51] of the
 * big ones; [52 + 2 b],[53 + 2 b] kilo-cycles and warp rounds of lane bucket b + 1 ([62],[63]: of the wide lane slot); [60] sweeps
 * summed over the lanes, [61] 32 x the warp's sweeps (their ratio is the lane occupancy of the lock-step position loop);
 * [11..14] cycles of stage A's phases and [15] its arena-ticks (diagnostic build only); [64..68] big islands: row passes, contact-sweeps,
 * contacts, bodies, sweeps (sums); [70..74] the same for the warp-slot islands. */
int ps_get_island_census(ps_handle h, int64_t* out /*[128]*/, int reset);
/* Per-kernel device time, measured with CUDA events on the engine's stream while enabled. ps_get_kernel_ms waits for the
 * stream, returns the accumulated milliseconds and launch counts of 0 = sense, 1 = think, 2 = physics (World.step, all stages),
 * 3 = other, 4 / 5 / 6 = the three stages of World.step (collide + islands, island workers, broad phase + TOI), 7 unused,
 * 8 / 9 / 10 = the big-island, warp-per-island and lane-batch kernels of stage B, each from the end of stage A (they run side by
 * side: the longest is what the tick waits for), 11 = the wide-lane-slot kernel (same convention); and clears the accumulators. */
int ps_profile(ps_handle h, int enable);
int ps_get_kernel_ms(ps_handle h, double* ms /*[12]*/, int64_t* counts /*[12]*/);
int     ps_num_groups(ps_handle h);        /* arena groups (each runs its kernel sequence on its own stream) */
int64_t ps_kernel_launches(ps_handle h);   /* kernels launched by this handle so far */
void*   ps_stream(ps_handle h);            /* the cudaStream_t the engine launches on */

#ifdef __cplusplus
}
#endif
#endif /* PUCKSWARM_B200_H */
