package puckswarm.b200;

import java.io.BufferedWriter;
import java.io.File;
import java.io.FileWriter;
import java.io.IOException;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.util.Properties;

/**
 * Drop-in for arena.Arena on the per-tick path (src/arena/Arena.java:45,183-267,310-316,392-411) backed by the CUDA
 * engine. It stands for a BATCH of arenas: arena i is the reference's `new Experiment(seed_i)` + `new Arena(world, null,
 * false)`. RunOffline keeps its loop (src/arena/RunOffline.java:72-95) but the three calls per think interval
 * (arena.step, arena.coast x2) and the three world.step calls collapse into one native call.
 *
 * NOT COMPILED in the build image of this repository (no JDK there); INTEGRATION.md has the build line. The layout
 * of {@link Config#toBuffer()} is checked at run time against ps_struct_size(0) by NativeEngine.
 */
public final class ArenaB200 {
    public static final int STORAGE_INTERVAL = 100;   // Arena.java:39
    private static final String[] COLOURS = { "RED", "GREEN", "MAGENTA", "CYAN", "ORANGE", "GRAY", "YELLOW", "PINK" };   // SensedType.java:9-23
    private final NativeEngine engine;
    private final Config cfg;
    private final long[] seeds;
    private final int nArenas, nRobots, nPucksSpawned, stepInterval;
    private int stepCount = 0;
    private final ByteBuffer pose;

    public ArenaB200(Config cfg, long[] seeds, Calibration calib, int device) {
        this.cfg = cfg; this.seeds = seeds.clone();
        this.nArenas = seeds.length; this.nRobots = cfg.nRobots; this.stepInterval = cfg.stepInterval;
        this.nPucksSpawned = cfg.nPucks / cfg.nPuckTypes * cfg.nPuckTypes;       // Arena.java:91-95
        cfg.nArenas = nArenas;
        engine = new NativeEngine(cfg.toBuffer(), device);
        engine.loadCalibration(calib.xrYr, calib.xy, calib.flags, calib.n, calib.width, calib.height);
        engine.reset(seeds);
        pose = direct(nArenas * nRobots * 3 * 4);
    }

    /** Arena.step(true, true, true, false, 0, false): perturbations, sense, think (on the device) and move; the World.step
     *  calls of the whole think interval follow in the same native call. */
    public void step(boolean allowThinking, boolean allowForwards, boolean allowTurning, boolean forcedMarch, int forcedTurn, boolean showCamera) {
        if (!allowThinking || !allowForwards || !allowTurning || forcedMarch || forcedTurn != 0)
            throw new UnsupportedOperationException("only Arena.step(true,true,true,false,0,false) is on the accelerated path");
        engine.step(stepInterval);
        stepCount++;
    }
    /** The last think step of an episode: RunOffline runs ONE World.step after it before its stop check (RunOffline.java:75-94). */
    public void lastStep() { engine.step(1); stepCount++; }
    /** Arena.coast: the coast ticks are part of step(); kept so that RunOffline's loop compiles unchanged. */
    public void coast(boolean allowThinking, boolean allowForwards, boolean allowTurning) { }
    public int getStepCount() { return stepCount; }
    /** Arena.getWorld (Arena.java:396-398): the engine handle stands for the JBox2D World of every arena of the batch. */
    public NativeEngine getWorld() { return engine; }
    /** Arena.getEnclosure (Arena.java:400-402): the rounded rectangle of RoundedRectangleEnclosure.java:37-44 at this scale. */
    public Enclosure getEnclosure() { return new Enclosure(cfg.enclosureScale); }
    /** x, y, theta of robot r of arena a as of the latest sense (SimAPS.getPose, SimAPS.java:14-18). */
    public float[] getRobotPose(int a, int r) {
        engine.getObservations(null, null, null, pose);
        int o = (a * nRobots + r) * 3 * 4;
        return new float[] { pose.getFloat(o), pose.getFloat(o + 4), pose.getFloat(o + 8) };
    }
    public void dispose() { engine.close(); }

    /** The files Arena.step and the controllers store every STORAGE_INTERVAL think steps (Arena.java:223-243,
     *  CacheConsController.java:313-333,531-551, BHDController.java:103-123), from the record the engine kept for the latest
     *  storage step, under outputDir/code/seed/ -- what analysis.Analyze reads. Call after a step() whose stepCount was a
     *  multiple of STORAGE_INTERVAL. */
    public void saveSnapshot(String outputDir, String code) throws IOException {
        int P = nPucksSpawned, R = nRobots;
        ByteBuffer step = direct(nArenas * 4), rp = direct(nArenas * R * 3 * 4), pk = direct(Math.max(1, nArenas * P * 2) * 4);
        ByteBuffer cache = direct(nArenas * R * 8 * 2 * 4), counts = direct(nArenas * R * 8 * 4);
        engine.getSnapshot(step, rp, pk, cache, counts);
        int perType = Math.max(1, P / Math.max(1, cfg.nPuckTypes));
        int nStates = cfg.controller == Config.CTRL_CACHECONS ? 7 : (cfg.controller == Config.CTRL_BHD ? 4 : 0);
        for (int a = 0; a < nArenas; a++) {
            int sc = step.getInt(a * 4);
            if (sc < 0) continue;
            File dir = new File(outputDir + File.separatorChar + code + File.separatorChar + seeds[a]);
            dir.mkdirs();
            String base = dir.getPath() + File.separatorChar + "step" + String.format("%07d", sc);
            try (BufferedWriter out = new BufferedWriter(new FileWriter(base + "_robots.txt"))) {      // PoseList.save: doubles
                for (int r = 0; r < R; r++) {
                    int o = ((a * R + r) * 3) * 4;
                    out.write((double) rp.getFloat(o) + ", " + (double) rp.getFloat(o + 4) + ", " + (double) rp.getFloat(o + 8) + "\n");
                }
            }
            for (int k = 0; k < cfg.nPuckTypes; k++) {
                int lo = k * perType, hi = Math.min(P, (k + 1) * perType);
                if (hi <= lo) continue;
                try (BufferedWriter out = new BufferedWriter(new FileWriter(base + "_" + COLOURS[k] + "_pucks.txt"))) {   // PositionList.save: floats
                    for (int p = lo; p < hi; p++) out.write(pk.getFloat(((a * P + p) * 2) * 4) + ", " + pk.getFloat(((a * P + p) * 2 + 1) * 4) + "\n");
                }
            }
            for (int r = 0; r < R && cfg.controller != Config.CTRL_EXTERNAL; r++) {
                if (cfg.controller == Config.CTRL_CACHECONS)
                    try (BufferedWriter out = new BufferedWriter(new FileWriter(base + "_R" + r + "_cachePoints.txt"))) {
                        for (int k = 0; k < 8; k++) {
                            int o = (((a * R + r) * 8 + k) * 2) * 4;
                            out.write(cache.getFloat(o) + ", " + cache.getFloat(o + 4) + "\n");      // Float.NaN prints "NaN"
                        }
                    }
                if (nStates > 0)
                    try (BufferedWriter out = new BufferedWriter(new FileWriter(base + "_R" + r + "_stateCounts.txt"))) {  // FileUtils.saveArray
                        for (int k = 0; k < nStates; k++) out.write(counts.getInt(((a * R + r) * 8 + k) * 4) + "\n");
                    }
            }
        }
    }

    static ByteBuffer direct(int bytes) { return ByteBuffer.allocateDirect(bytes).order(ByteOrder.nativeOrder()); }

    public static final class Enclosure {
        public final float WIDTH, HEIGHT, R;
        Enclosure(float scale) { WIDTH = 187f * scale; HEIGHT = 187f * scale; R = 15.5f * scale; }
    }

    /** The java.util.Properties keys of SURVEY App. C.8 as a ps_config image: 44 four-byte fields in header order
     *  (include/puckswarm_b200.h: struct ps_config). Missing key = the reference's default; malformed number =
     *  NumberFormatException, like Experiment.getProperty (Experiment.java:76-90). */
    public static final class Config {
        public static final int ABI_VERSION = 3;
        public static final int CTRL_CACHECONS = 0, CTRL_BHD = 1, CTRL_EXTERNAL = 2, CTRL_PROBSEEK = 3;
        public int nArenas = 1, nRobots = 4, nPucks = 10, nPuckTypes = 2, nInformedRobots = 0, controller = CTRL_CACHECONS;
        public int rngMode = 0, stepInterval = 3, velIters = 3, posIters = 100;
        public float dt = 1f / 14f, enclosureScale = 1f;
        public int sincosLut = 1, toiEnabled = 1, maxContacts = 0;
        public float clusterDistanceThreshold = 1.5f;                                     // LocalMap.java:74 (a constant, not a property)
        public Properties properties = new Properties();

        public Config() { }
        /** common.properties overlaid by code.properties, as Experiment does (Experiment.java:26-33). */
        public Config(Properties p) {
            properties = p;
            nRobots = geti("Arena.nRobots", 4); nPucks = geti("Arena.nPucks", 10); nPuckTypes = geti("Arena.nPuckTypes", 2);
            nInformedRobots = geti("Arena.nInformedRobots", 0); enclosureScale = getf("RoundedRectangleEnclosure.scale", 1f);
            String c = p.getProperty("controllerType", "CacheCons");
            if (c.equals("CacheCons")) controller = CTRL_CACHECONS; else if (c.equals("BHD")) controller = CTRL_BHD;
            else if (c.equals("ProbSeek")) controller = CTRL_PROBSEEK; else controller = CTRL_EXTERNAL;   // any other controller stays in Java
        }
        private int geti(String k, int d) { String v = properties.getProperty(k); return v == null ? d : Integer.parseInt(v.trim()); }
        private float getf(String k, float d) { String v = properties.getProperty(k); return v == null ? d : Float.parseFloat(v.trim()); }
        private int getb(String k, boolean d) { String v = properties.getProperty(k); return (v == null ? d : Boolean.parseBoolean(v.trim())) ? 1 : 0; }
        private int gete(String k, int d, String... names) {
            String v = properties.getProperty(k);
            if (v == null) return d;
            for (int i = 0; i < names.length; i++) if (names[i].equals(v.trim())) return i;
            throw new IllegalArgumentException(k + ": unknown value " + v);                  // Enum.valueOf
        }
        public ByteBuffer toBuffer() {
            ByteBuffer b = direct(44 * 4);
            // ProbSeek reads K1 .. ST_DEV under its own class prefix (ProbSeekController.java:98-121); they travel in the cc_* fields
            String cc = controller == CTRL_PROBSEEK ? "ProbSeekController." : "CacheConsController.";
            b.putInt(ABI_VERSION).putInt(nArenas).putInt(nRobots).putInt(nPucks).putInt(nPuckTypes).putInt(nInformedRobots);
            b.putInt(controller).putInt(rngMode).putInt(stepInterval).putInt(velIters).putInt(posIters);
            b.putFloat(dt).putFloat(enclosureScale).putInt(sincosLut).putInt(toiEnabled).putInt(maxContacts);
            b.putFloat(clusterDistanceThreshold);
            b.putFloat(getf("LocalMap.CARRY_THRESHOLD_LO", 0.1f)).putFloat(getf("LocalMap.CARRY_THRESHOLD_HI", 0.4f));
            b.putFloat(getf("VFHPlus.SAFETY_DISTANCE_GAMMA", 2f)).putFloat(getf("VFHPlus.SAFETY_DISTANCE_MASK", 5f));
            b.putFloat(getf("VFHPlus.TAU_LO", 0.7f)).putFloat(getf("VFHPlus.TAU_HI", 0.85f));
            b.putFloat(getf(cc + "K1", 1f)).putFloat(getf(cc + "K2", 8f));
            b.putInt(geti(cc + "PICK_UP_TIME", 20)).putInt(geti(cc + "PUSH_IN_TIME", 1)).putInt(geti(cc + "BACKUP_TIME", 5));
            b.putFloat(getf(cc + "PREDICTION_DISTANCE_SQD", 100f)).putFloat(getf(cc + "ST_DEV", 0.5f));
            b.putInt(geti("CacheConsController.EXILE_TIME", 20)).putInt(geti("CacheConsController.SPIT_OUT_TIME", 10));
            b.putInt(gete("CacheConsController.CACHE_SELECTION_OPTION", 2, "PROB_ABSOLUTE", "PROB_RELATIVE", "RELATIVE", "RELATIVE_FORGET"));
            b.putFloat(getf("CacheConsController.K_ABSOLUTE", 40f)).putFloat(getf("CacheConsController.K_RELATIVE", 8f));
            b.putFloat(getf("CacheConsController.FORGET_PROB", 0.001f));
            b.putInt(gete("CacheConsController.CACHE_SEPARATION_OPTION", 3, "IGNORE", "NULLIFY_OLD", "ACCEPT_ISOLATED", "ACCEPT_LARGER"));
            b.putFloat(getf("CacheConsController.HOME_SEPARATION_DISTANCE", 50f)).putInt(getb("CacheConsController.IGNORE_PUCKS", false));
            b.putInt(geti("BHDController.PUSH_IN_TIME", 1)).putInt(geti("BHDController.BACKUP_TIME", 5));
            b.putInt(geti("ProbSeekController.PLACEMENT_TIME", 40));
            b.putInt(geti("Arena.puckShiftStep", -1)).putInt(geti("Arena.puckShuffleStep", -1));
            if (b.position() != 44 * 4) throw new IllegalStateException("ps_config image has " + b.position() + " bytes, expected " + 44 * 4);
            b.rewind();
            return b;
        }
    }
    public static final class Calibration { public ByteBuffer xrYr, xy, flags; public int n, width, height; }
}
