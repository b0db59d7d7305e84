/*
 * JNI glue between puckswarm.b200.NativeEngine and the C-ABI of include/puckswarm_b200.h.
 * NOT COMPILED in this repository's build image (no JDK / jni.h there). Build line (on a box with a JDK):
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include \
 *       -o libpuckswarm_b200_jni.so puckswarm_jni.c -L../../puckswarm_b200/csrc -lpuckswarm_b200
 */
#include <jni.h>
#include <string.h>
#include "puckswarm_b200.h"

#define BUF(env, b) ((b) ? (*(env))->GetDirectBufferAddress((env), (b)) : NULL)
#define FN(name) Java_puckswarm_b200_NativeEngine_##name

JNIEXPORT jstring JNICALL FN(lastError)(JNIEnv* env, jclass c) { return (*env)->NewStringUTF(env, ps_last_error()); }

JNIEXPORT jint JNICALL FN(psCreate)(JNIEnv* env, jclass c, jobject cfg, jint device, jlongArray out) {
    ps_handle h = NULL;
    int rc = ps_create((const ps_config*)BUF(env, cfg), device, &h);
    jlong v = (jlong)(intptr_t)h;
    (*env)->SetLongArrayRegion(env, out, 0, 1, &v);
    return rc;
}
JNIEXPORT jint JNICALL FN(psDestroy)(JNIEnv* env, jclass c, jlong h) { return ps_destroy((ps_handle)(intptr_t)h); }
JNIEXPORT jint JNICALL FN(psLoadCalibration)(JNIEnv* env, jclass c, jlong h, jobject xr, jobject xy, jobject fl, jint n, jint w, jint hh) {
    return ps_load_calibration((ps_handle)(intptr_t)h, (const float*)BUF(env, xr), (const uint16_t*)BUF(env, xy), (const uint8_t*)BUF(env, fl), n, w, hh);
}
JNIEXPORT jint JNICALL FN(psReset)(JNIEnv* env, jclass c, jlong h, jlongArray seeds) {
    jlong* s = (*env)->GetLongArrayElements(env, seeds, NULL);
    int rc = ps_reset((ps_handle)(intptr_t)h, (const int64_t*)s);
    (*env)->ReleaseLongArrayElements(env, seeds, s, JNI_ABORT);
    return rc;
}
JNIEXPORT jint JNICALL FN(psStep)(JNIEnv* env, jclass c, jlong h, jint n) { return ps_step((ps_handle)(intptr_t)h, n); }
JNIEXPORT jint JNICALL FN(psSync)(JNIEnv* env, jclass c, jlong h) { return ps_sync((ps_handle)(intptr_t)h); }
JNIEXPORT jint JNICALL FN(psSense)(JNIEnv* env, jclass c, jlong h) { return ps_sense((ps_handle)(intptr_t)h); }
JNIEXPORT jint JNICALL FN(psStepExternal)(JNIEnv* env, jclass c, jlong h, jfloatArray f, jfloatArray t) {
    jfloat* pf = (*env)->GetFloatArrayElements(env, f, NULL);
    jfloat* pt = (*env)->GetFloatArrayElements(env, t, NULL);
    int rc = ps_step_external((ps_handle)(intptr_t)h, pf, pt);   /* copies before returning */
    (*env)->ReleaseFloatArrayElements(env, f, pf, JNI_ABORT);
    (*env)->ReleaseFloatArrayElements(env, t, pt, JNI_ABORT);
    return rc;
}
JNIEXPORT jint JNICALL FN(psGetObservations)(JNIEnv* env, jclass c, jlong h, jobject cam, jobject occ, jobject lm, jobject pose) {
    ps_obs_view v;
    v.camera = (uint8_t*)BUF(env, cam);
    v.occupancy = (uint8_t*)BUF(env, occ);
    v.localmap = (ps_localmap*)BUF(env, lm);
    v.pose = (float*)BUF(env, pose);
    return ps_get_observations((ps_handle)(intptr_t)h, &v);
}
JNIEXPORT jint JNICALL FN(psGetBodies)(JNIEnv* env, jclass c, jlong h, jobject body, jobject arena) {
    ps_state_view v;
    memset(&v, 0, sizeof(v));
    v.body = (float*)BUF(env, body);
    v.arena = (ps_arena_state*)BUF(env, arena);
    return ps_get_state((ps_handle)(intptr_t)h, &v);
}
JNIEXPORT jint JNICALL FN(psComputeStats)(JNIEnv* env, jclass c, jlong h, jfloat thr, jobject stats) {
    return ps_compute_stats((ps_handle)(intptr_t)h, thr, (ps_stats*)BUF(env, stats));
}
JNIEXPORT jint JNICALL FN(psGetStats)(JNIEnv* env, jclass c, jlong h, jfloat thr, jobject stats, jlong comm) {
    return ps_get_stats((ps_handle)(intptr_t)h, thr, (ps_stats*)BUF(env, stats), (void*)(intptr_t)comm);
}
JNIEXPORT jint JNICALL FN(psGetSnapshot)(JNIEnv* env, jclass c, jlong h, jobject step, jobject pose, jobject puck, jobject cache, jobject counts) {
    ps_snapshot_view v;
    v.step_count = (int32_t*)BUF(env, step);
    v.robot_pose = (float*)BUF(env, pose);
    v.puck_xy = (float*)BUF(env, puck);
    v.cache_xy = (float*)BUF(env, cache);
    v.state_counts = (int32_t*)BUF(env, counts);
    return ps_get_snapshot((ps_handle)(intptr_t)h, &v);
}
JNIEXPORT jint JNICALL FN(psGetAnalysis)(JNIEnv* env, jclass c, jlong h, jobject out) {
    return ps_get_analysis((ps_handle)(intptr_t)h, (ps_arena_analysis*)BUF(env, out));
}
JNIEXPORT jint JNICALL FN(psGetErrorFlags)(JNIEnv* env, jclass c, jlong h, jintArray out) {
    int32_t v[2] = {0, 0};
    int rc = ps_get_error_flags((ps_handle)(intptr_t)h, &v[0], &v[1]);
    jint j[2] = {v[0], v[1]};
    (*env)->SetIntArrayRegion(env, out, 0, 2, j);
    return rc;
}
JNIEXPORT jint JNICALL FN(psStructSize)(JNIEnv* env, jclass c, jint which) { return ps_struct_size(which); }
