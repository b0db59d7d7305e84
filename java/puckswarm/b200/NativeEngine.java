package puckswarm.b200;

import java.nio.ByteBuffer;

/**
 * JNI face of libpuckswarm_b200.so (include/puckswarm_b200.h). One instance = one ps_handle = one batch of arenas on
 * one GPU, confined to one Java thread. Every native method returns the C error code (0 = PS_OK); on failure
 * {@link #lastError()} holds the message. Direct ByteBuffers are used for the caller-allocated arrays of the C-ABI so
 * that no copy happens on the Java side.
 *
 * NOT COMPILED in the build image of this repository (no JDK there); see INTEGRATION.md for the build line.
 */
public final class NativeEngine implements AutoCloseable {
    static { System.loadLibrary("puckswarm_b200_jni"); }

    private long handle;   // ps_handle

    /** cfg: a ps_config image laid out as in the header (use {@link Config#toBuffer()}). */
    public NativeEngine(ByteBuffer cfg, int device) {
        if (psStructSize(0) != cfg.capacity())
            throw new IllegalStateException("ps_config layout mismatch: library " + psStructSize(0) + " bytes, Java image " + cfg.capacity());
        long[] out = new long[1];
        check(psCreate(cfg, device, out));
        handle = out[0];
    }

    public void loadCalibration(ByteBuffer xrYr, ByteBuffer xy, ByteBuffer flags, int n, int w, int h) { check(psLoadCalibration(handle, xrYr, xy, flags, n, w, h)); }
    public void reset(long[] seeds) { check(psReset(handle, seeds)); }
    /** n ticks of RunOffline.step(): {Arena.step | Arena.coast} + World.step(1/14f, 3, 100). Asynchronous. */
    public void step(int nTicks) { check(psStep(handle, nTicks)); }
    public void sync() { check(psSync(handle)); }
    /** Host-controller path: sense now, then one think interval with the given commands [nArenas][nRobots]. */
    public void sense() { check(psSense(handle)); }
    public void stepExternal(float[] forwards, float[] torque) { check(psStepExternal(handle, forwards, torque)); }
    /** poses [nArenas][nRobots][3], local maps and occupancy grids into direct buffers (null = skip). */
    public void getObservations(ByteBuffer camera, ByteBuffer occupancy, ByteBuffer localmap, ByteBuffer pose) { check(psGetObservations(handle, camera, occupancy, localmap, pose)); }
    public void getBodies(ByteBuffer body, ByteBuffer arena) { check(psGetBodies(handle, body, arena)); }
    public void computeStats(float clusterThreshold, ByteBuffer stats) { check(psComputeStats(handle, clusterThreshold, stats)); }
    /** ps_stats reduced over the ranks of an NCCL communicator (ncclComm_t as a long; 0 = this process only). */
    public void getStats(float clusterThreshold, ByteBuffer stats, long ncclComm) { check(psGetStats(handle, clusterThreshold, stats, ncclComm)); }
    /** The reference's periodic record of the latest storage step (ps_snapshot_view; null = skip a component). */
    public void getSnapshot(ByteBuffer stepCount, ByteBuffer robotPose, ByteBuffer puckXy, ByteBuffer cacheXy, ByteBuffer stateCounts) {
        check(psGetSnapshot(handle, stepCount, robotPose, puckXy, cacheXy, stateCounts));
    }
    /** ps_arena_analysis[nArenas]: Analyze's stepsToCompletion / timeWeightedCompletion accumulators. */
    public void getAnalysis(ByteBuffer out) { check(psGetAnalysis(handle, out)); }
    /** {OR of the arenas' capacity flags, flagged arenas}; sync() throws on the same condition. */
    public int[] getErrorFlags() { int[] o = new int[2]; check(psGetErrorFlags(handle, o)); return o; }
    @Override public void close() { if (handle != 0) { psDestroy(handle); handle = 0; } }

    public static native String lastError();
    private static void check(int rc) { if (rc != 0) throw new IllegalStateException("puckswarm_b200 error " + rc + ": " + lastError()); }

    private static native int psCreate(ByteBuffer cfg, int device, long[] outHandle);
    private static native int psDestroy(long h);
    private static native int psLoadCalibration(long h, ByteBuffer xrYr, ByteBuffer xy, ByteBuffer flags, int n, int w, int hgt);
    private static native int psReset(long h, long[] seeds);
    private static native int psStep(long h, int nTicks);
    private static native int psSync(long h);
    private static native int psSense(long h);
    private static native int psStepExternal(long h, float[] forwards, float[] torque);
    private static native int psGetObservations(long h, ByteBuffer camera, ByteBuffer occupancy, ByteBuffer localmap, ByteBuffer pose);
    private static native int psGetBodies(long h, ByteBuffer body, ByteBuffer arena);
    private static native int psComputeStats(long h, float clusterThreshold, ByteBuffer stats);
    private static native int psGetStats(long h, float clusterThreshold, ByteBuffer stats, long ncclComm);
    private static native int psGetSnapshot(long h, ByteBuffer stepCount, ByteBuffer robotPose, ByteBuffer puckXy, ByteBuffer cacheXy, ByteBuffer stateCounts);
    private static native int psGetAnalysis(long h, ByteBuffer out);
    private static native int psGetErrorFlags(long h, int[] out);
    private static native int psStructSize(int which);
}
