/*
 * jni_min.h -- minimal, specification-conformant subset of <jni.h>.
 *
 * There is no JDK in the build image, so the JNI shim (beast-mcmc_b200/csrc/jni_shim.cpp) is
 * compiled against this header instead.  It is written from the JNI specification: the primitive
 * typedefs and the POSITIONS of the entries of JNINativeInterface_ (the function table a JVM hands
 * to native code) are fixed by the specification, so a library built against this header is
 * binary-compatible with a real JVM.  Only the entries the shim uses get typed accessors; the
 * table itself is declared as an array of untyped slots.
 */
#ifndef B200_JNI_MIN_H
#define B200_JNI_MIN_H

#include <stdarg.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JNIEXPORT __attribute__((visibility("default")))
#define JNIIMPORT
#define JNICALL

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef uint16_t jchar;
typedef int16_t jshort;
typedef float jfloat;
typedef double jdouble;
typedef jint jsize;

struct _jobject;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jobjectArray;
typedef jarray jintArray;
typedef jarray jdoubleArray;
typedef jobject jthrowable;
struct _jmethodID;
typedef struct _jmethodID* jmethodID;

#define JNI_FALSE 0
#define JNI_TRUE 1
#define JNI_OK 0
#define JNI_COMMIT 1
#define JNI_ABORT 2
#define JNI_VERSION_1_6 0x00010006

/* indices into the JNI function table (JNI specification, "Interface Function Table") */
enum {
    JNI_IDX_GetVersion = 4,
    JNI_IDX_FindClass = 6,
    JNI_IDX_ExceptionClear = 17,
    JNI_IDX_DeleteLocalRef = 23,
    JNI_IDX_NewObject = 28,
    JNI_IDX_GetObjectClass = 31,
    JNI_IDX_GetMethodID = 33,
    JNI_IDX_CallVoidMethod = 61,
    JNI_IDX_NewStringUTF = 167,
    JNI_IDX_GetArrayLength = 171,
    JNI_IDX_NewObjectArray = 172,
    JNI_IDX_SetObjectArrayElement = 174,
    JNI_IDX_GetIntArrayElements = 187,
    JNI_IDX_GetDoubleArrayElements = 190,
    JNI_IDX_ReleaseIntArrayElements = 195,
    JNI_IDX_ReleaseDoubleArrayElements = 198,
    JNI_IDX_GetPrimitiveArrayCritical = 222,
    JNI_IDX_ReleasePrimitiveArrayCritical = 223,
    JNI_IDX_ExceptionCheck = 228,
    JNI_TABLE_SLOTS = 240
};

struct JNINativeInterface_ {
    void* slot[JNI_TABLE_SLOTS];
};

#ifdef __cplusplus
}  /* extern "C" */
struct JNIEnv_;
typedef JNIEnv_ JNIEnv;
struct JNIEnv_ {
    const struct JNINativeInterface_* functions;

    template <typename F> F fn(int idx) const { return reinterpret_cast<F>(functions->slot[idx]); }

    jint GetVersion() { return fn<jint (*)(JNIEnv*)>(JNI_IDX_GetVersion)(this); }
    jclass FindClass(const char* name) { return fn<jclass (*)(JNIEnv*, const char*)>(JNI_IDX_FindClass)(this, name); }
    void ExceptionClear() { fn<void (*)(JNIEnv*)>(JNI_IDX_ExceptionClear)(this); }
    jboolean ExceptionCheck() { return fn<jboolean (*)(JNIEnv*)>(JNI_IDX_ExceptionCheck)(this); }
    void DeleteLocalRef(jobject o) { fn<void (*)(JNIEnv*, jobject)>(JNI_IDX_DeleteLocalRef)(this, o); }
    jmethodID GetMethodID(jclass c, const char* name, const char* sig) {
        return fn<jmethodID (*)(JNIEnv*, jclass, const char*, const char*)>(JNI_IDX_GetMethodID)(this, c, name, sig);
    }
    template <typename... Args> jobject NewObject(jclass c, jmethodID m, Args... args) {
        return fn<jobject (*)(JNIEnv*, jclass, jmethodID, ...)>(JNI_IDX_NewObject)(this, c, m, args...);
    }
    template <typename... Args> void CallVoidMethod(jobject o, jmethodID m, Args... args) {
        fn<void (*)(JNIEnv*, jobject, jmethodID, ...)>(JNI_IDX_CallVoidMethod)(this, o, m, args...);
    }
    jstring NewStringUTF(const char* s) { return fn<jstring (*)(JNIEnv*, const char*)>(JNI_IDX_NewStringUTF)(this, s); }
    jsize GetArrayLength(jarray a) { return fn<jsize (*)(JNIEnv*, jarray)>(JNI_IDX_GetArrayLength)(this, a); }
    jobjectArray NewObjectArray(jsize n, jclass c, jobject init) {
        return fn<jobjectArray (*)(JNIEnv*, jsize, jclass, jobject)>(JNI_IDX_NewObjectArray)(this, n, c, init);
    }
    void SetObjectArrayElement(jobjectArray a, jsize i, jobject v) {
        fn<void (*)(JNIEnv*, jobjectArray, jsize, jobject)>(JNI_IDX_SetObjectArrayElement)(this, a, i, v);
    }
    jint* GetIntArrayElements(jintArray a, jboolean* isCopy) {
        return fn<jint* (*)(JNIEnv*, jintArray, jboolean*)>(JNI_IDX_GetIntArrayElements)(this, a, isCopy);
    }
    jdouble* GetDoubleArrayElements(jdoubleArray a, jboolean* isCopy) {
        return fn<jdouble* (*)(JNIEnv*, jdoubleArray, jboolean*)>(JNI_IDX_GetDoubleArrayElements)(this, a, isCopy);
    }
    void ReleaseIntArrayElements(jintArray a, jint* e, jint mode) {
        fn<void (*)(JNIEnv*, jintArray, jint*, jint)>(JNI_IDX_ReleaseIntArrayElements)(this, a, e, mode);
    }
    void ReleaseDoubleArrayElements(jdoubleArray a, jdouble* e, jint mode) {
        fn<void (*)(JNIEnv*, jdoubleArray, jdouble*, jint)>(JNI_IDX_ReleaseDoubleArrayElements)(this, a, e, mode);
    }
    void* GetPrimitiveArrayCritical(jarray a, jboolean* isCopy) {
        return fn<void* (*)(JNIEnv*, jarray, jboolean*)>(JNI_IDX_GetPrimitiveArrayCritical)(this, a, isCopy);
    }
    void ReleasePrimitiveArrayCritical(jarray a, void* e, jint mode) {
        fn<void (*)(JNIEnv*, jarray, void*, jint)>(JNI_IDX_ReleasePrimitiveArrayCritical)(this, a, e, mode);
    }
};
#else
typedef const struct JNINativeInterface_* JNIEnv;
#endif

#endif /* B200_JNI_MIN_H */
