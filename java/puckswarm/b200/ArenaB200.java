package puckswarm.b200;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;

/**
 * Drop-in for arena.Arena on the per-tick path (src/arena/Arena.java:45,183-267,310-316,392-411) backed by the CUDA
 * engine. It stands for a BATCH of arenas: arena i is the reference's `new Experiment(seed_i)` + `new Arena(world, null,
 * false)`. RunOffline keeps its loop (src/arena/RunOffline.java:72-95) but the three calls per think interval
 * (arena.step, arena.coast x2) and the three world.step calls collapse into one native call.
 *
 * NOT COMPILED in the build image of this repository (no JDK there).
 */
public final class ArenaB200 {
    public static final int STORAGE_INTERVAL = 100;   // Arena.java:39
    private final NativeEngine engine;
    private final int nArenas, nRobots, stepInterval;
    private int stepCount = 0, ticksIntoInterval = 0;
    private final ByteBuffer pose;

    public ArenaB200(Config cfg, long[] seeds, Calibration calib, int device) {
        this.nArenas = cfg.nArenas; this.nRobots = cfg.nRobots; this.stepInterval = cfg.stepInterval;
        engine = new NativeEngine(cfg.toBuffer(), device);
        engine.loadCalibration(calib.xrYr, calib.xy, calib.flags, calib.n, calib.width, calib.height);
        engine.reset(seeds);
        pose = ByteBuffer.allocateDirect(nArenas * nRobots * 3 * 4).order(ByteOrder.nativeOrder());
    }

    /** Arena.step(true, true, true, false, 0, false): sense, think (on the device) and move; the World.step calls of
     *  the whole think interval follow in the same native call. */
    public void step(boolean allowThinking, boolean allowForwards, boolean allowTurning, boolean forcedMarch, int forcedTurn, boolean showCamera) {
        if (!allowThinking || !allowForwards || !allowTurning || forcedMarch || forcedTurn != 0)
            throw new UnsupportedOperationException("only Arena.step(true,true,true,false,0,false) is on the accelerated path");
        engine.step(stepInterval);
        stepCount++;
    }
    /** Arena.coast: the coast ticks are part of step(); kept so that RunOffline's loop compiles unchanged. */
    public void coast(boolean allowThinking, boolean allowForwards, boolean allowTurning) { }
    public int getStepCount() { return stepCount; }
    /** x, y, theta of robot r of arena a as Arena.java:231-233 would store them. */
    public float[] getRobotPose(int a, int r) {
        engine.getObservations(null, null, null, pose);
        int o = (a * nRobots + r) * 3 * 4;
        return new float[] { pose.getFloat(o), pose.getFloat(o + 4), pose.getFloat(o + 8) };
    }
    public void dispose() { engine.close(); }

    /** The java.util.Properties keys of SURVEY App. C.8 as a ps_config image. */
    public static final class Config {
        public int nArenas = 1, nRobots = 4, nPucks = 10, nPuckTypes = 2, stepInterval = 3;
        public java.util.Properties properties = new java.util.Properties();
        public ByteBuffer toBuffer() { throw new UnsupportedOperationException("fill the 44 fields of ps_config in header order; ps_struct_size(0) gives the size to check against"); }
    }
    public static final class Calibration { public ByteBuffer xrYr, xy, flags; public int n, width, height; }
}
