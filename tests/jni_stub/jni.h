/* TEST INFRASTRUCTURE ONLY — a declarations-only stand-in for the JDK's <jni.h>, so java/jni/b2c_jni.c can be
 * compile-checked in an image without a JDK.  It declares exactly the types and JNIEnv entries the forwarders use, with the
 * JNI specification's signatures; it is never linked into anything that runs. */
#ifndef JNI_STUB_H
#define JNI_STUB_H
#include <stdint.h>
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
typedef int32_t jint;
typedef int64_t jlong;
typedef int16_t jshort;
typedef float jfloat;
typedef jint jsize;
struct _jobject;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jlongArray;
typedef jarray jfloatArray;
typedef jarray jobjectArray;
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
    void* (*GetDirectBufferAddress)(JNIEnv* env, jobject buf);
    jstring (*NewStringUTF)(JNIEnv* env, const char* utf);
    jsize (*GetArrayLength)(JNIEnv* env, jarray array);
    jobject (*GetObjectArrayElement)(JNIEnv* env, jobjectArray array, jsize index);
    void (*GetIntArrayRegion)(JNIEnv* env, jintArray array, jsize start, jsize len, jint* buf);
    void (*SetIntArrayRegion)(JNIEnv* env, jintArray array, jsize start, jsize len, const jint* buf);
    void (*SetLongArrayRegion)(JNIEnv* env, jlongArray array, jsize start, jsize len, const jlong* buf);
    void (*SetFloatArrayRegion)(JNIEnv* env, jfloatArray array, jsize start, jsize len, const jfloat* buf);
};
#endif
