// jni_shim.cpp -- libhmsbeagle-jni.so: the 47 `native` methods of beagle.BeagleJNIWrapper
// (lib/beagle.jar of the reference; descriptors in tests/golden/beagle_jar_abi.json) forwarded 1:1 to
// the C ABI of libhmsbeagle.so.  BEAST loads this library with System.loadLibrary("hmsbeagle-jni")
// (disassembly of BeagleJNIWrapper.loadBeagleLibrary; -Djava.library.path in tests/test.sh:10).
// Natives are INSTANCE methods of the singleton BeagleJNIWrapper.INSTANCE, hence the jobject second
// argument.  Java arrays are only valid for the duration of a call: inputs are released with
// JNI_ABORT, outputs with mode 0 (copy back).  Nullable arrays (derivative indices on the hot path,
// HomogenousSubstitutionModelDelegate.java:258-263; resourceList) map to nullptr.
#include "jni_min.h"
#include "libhmsbeagle_b200.h"

#include <vector>

namespace {

struct IntIn {          // read-only int[] (may be null)
    JNIEnv* env; jintArray arr; jint* p;
    IntIn(JNIEnv* e, jintArray a) : env(e), arr(a), p(a ? e->GetIntArrayElements(a, nullptr) : nullptr) {}
    ~IntIn() { if (p) env->ReleaseIntArrayElements(arr, p, JNI_ABORT); }
    operator const int*() const { return reinterpret_cast<const int*>(p); }
    int* mut() { return reinterpret_cast<int*>(p); }
};
struct IntOut {
    JNIEnv* env; jintArray arr; jint* p;
    IntOut(JNIEnv* e, jintArray a) : env(e), arr(a), p(a ? e->GetIntArrayElements(a, nullptr) : nullptr) {}
    ~IntOut() { if (p) env->ReleaseIntArrayElements(arr, p, 0); }
    operator int*() { return reinterpret_cast<int*>(p); }
};
struct DblIn {
    JNIEnv* env; jdoubleArray arr; jdouble* p;
    DblIn(JNIEnv* e, jdoubleArray a) : env(e), arr(a), p(a ? e->GetDoubleArrayElements(a, nullptr) : nullptr) {}
    ~DblIn() { if (p) env->ReleaseDoubleArrayElements(arr, p, JNI_ABORT); }
    operator const double*() const { return p; }
};
struct DblOut {
    JNIEnv* env; jdoubleArray arr; jdouble* p;
    DblOut(JNIEnv* e, jdoubleArray a) : env(e), arr(a), p(a ? e->GetDoubleArrayElements(a, nullptr) : nullptr) {}
    ~DblOut() { if (p) env->ReleaseDoubleArrayElements(arr, p, 0); }
    operator double*() { return p; }
};

// Read-only view of the arrays on the per-step path (operation lists, matrix indices, branch lengths, scale indices):
// Get<Type>ArrayElements COPIES on HotSpot; GetPrimitiveArrayCritical hands out the array itself (the collector is
// merely held off for the call).  The engine copies what it needs into its pinned staging ring and enqueues -- no JNI
// call and no wait on another Java thread happens inside the critical region, as the JNI specification demands.
template <typename T>
struct CritIn {
    JNIEnv* env; jarray arr; void* p;
    CritIn(JNIEnv* e, jarray a) : env(e), arr(a), p(a ? e->GetPrimitiveArrayCritical(a, nullptr) : nullptr) {}
    ~CritIn() { if (p) env->ReleasePrimitiveArrayCritical(arr, p, JNI_ABORT); }
    operator const T*() const { return static_cast<const T*>(p); }
};

}  // namespace

#define NATIVE(ret, name) extern "C" JNIEXPORT ret JNICALL Java_beagle_BeagleJNIWrapper_##name

NATIVE(jstring, getVersion)(JNIEnv* env, jobject) { return env->NewStringUTF(beagleGetVersion()); }
NATIVE(jstring, getCitation)(JNIEnv* env, jobject) { return env->NewStringUTF(beagleGetCitation()); }

NATIVE(jobjectArray, getResourceList)(JNIEnv* env, jobject) {
    BeagleResourceList* rl = beagleGetResourceList();
    if (rl == nullptr) return nullptr;
    jclass cls = env->FindClass("beagle/ResourceDetails");
    if (cls == nullptr) return nullptr;
    jmethodID ctor = env->GetMethodID(cls, "<init>", "(I)V");
    jmethodID setName = env->GetMethodID(cls, "setName", "(Ljava/lang/String;)V");
    jmethodID setDesc = env->GetMethodID(cls, "setDescription", "(Ljava/lang/String;)V");
    jmethodID setFlags = env->GetMethodID(cls, "setFlags", "(J)V");
    if (!ctor || !setName || !setDesc || !setFlags) return nullptr;
    jobjectArray out = env->NewObjectArray(rl->length, cls, nullptr);
    for (int i = 0; i < rl->length; ++i) {
        jobject r = env->NewObject(cls, ctor, (jint)i);
        jstring n = env->NewStringUTF(rl->list[i].name);
        jstring d = env->NewStringUTF(rl->list[i].description);
        env->CallVoidMethod(r, setName, n);
        env->CallVoidMethod(r, setDesc, d);
        env->CallVoidMethod(r, setFlags, (jlong)rl->list[i].supportFlags);
        env->SetObjectArrayElement(out, i, r);
        env->DeleteLocalRef(n); env->DeleteLocalRef(d); env->DeleteLocalRef(r);
    }
    return out;
}

NATIVE(jobjectArray, getBenchmarkedResourceList)(JNIEnv* env, jobject, jint tipCount, jint compactBufferCount,
                                                 jint stateCount, jint patternCount, jint categoryCount,
                                                 jintArray resourceList, jint resourceCount, jlong preferenceFlags,
                                                 jlong requirementFlags, jint eigenModelCount, jint partitionCount,
                                                 jint calculateDerivatives, jlong benchmarkFlags) {
    IntIn res(env, resourceList);
    BeagleBenchmarkedResourceList* bl = beagleGetBenchmarkedResourceList(
        tipCount, compactBufferCount, stateCount, patternCount, categoryCount, res.mut(), res.p ? resourceCount : 0,
        (long)preferenceFlags, (long)requirementFlags, eigenModelCount, partitionCount, calculateDerivatives,
        (long)benchmarkFlags);
    if (bl == nullptr) return nullptr;
    jclass cls = env->FindClass("beagle/BenchmarkedResourceDetails");
    if (cls == nullptr) return nullptr;
    jmethodID ctor = env->GetMethodID(cls, "<init>", "(I)V");
    jmethodID setResourceNumber = env->GetMethodID(cls, "setResourceNumber", "(I)V");
    jmethodID setName = env->GetMethodID(cls, "setName", "(Ljava/lang/String;)V");
    jmethodID setDesc = env->GetMethodID(cls, "setDescription", "(Ljava/lang/String;)V");
    jmethodID setSupport = env->GetMethodID(cls, "setSupportFlags", "(J)V");
    jmethodID setRequired = env->GetMethodID(cls, "setRequiredFlags", "(J)V");
    jmethodID setReturnCode = env->GetMethodID(cls, "setReturnCode", "(I)V");
    jmethodID setImplName = env->GetMethodID(cls, "setImplName", "(Ljava/lang/String;)V");
    jmethodID setBenched = env->GetMethodID(cls, "setBenchedFlags", "(J)V");
    jmethodID setResult = env->GetMethodID(cls, "setBenchmarkResult", "(D)V");
    jmethodID setRatio = env->GetMethodID(cls, "setPerformanceRatio", "(D)V");
    if (!ctor || !setResourceNumber || !setName || !setDesc || !setSupport || !setRequired || !setReturnCode ||
        !setImplName || !setBenched || !setResult || !setRatio)
        return nullptr;
    jobjectArray out = env->NewObjectArray(bl->length, cls, nullptr);
    for (int i = 0; i < bl->length; ++i) {
        const BeagleBenchmarkedResource& b = bl->list[i];
        jobject r = env->NewObject(cls, ctor, (jint)i);
        jstring n = env->NewStringUTF(b.name), d = env->NewStringUTF(b.description), im = env->NewStringUTF(b.implName);
        env->CallVoidMethod(r, setResourceNumber, (jint)b.number);
        env->CallVoidMethod(r, setName, n);
        env->CallVoidMethod(r, setDesc, d);
        env->CallVoidMethod(r, setSupport, (jlong)b.supportFlags);
        env->CallVoidMethod(r, setRequired, (jlong)b.requiredFlags);
        env->CallVoidMethod(r, setReturnCode, (jint)b.returnCode);
        env->CallVoidMethod(r, setImplName, im);
        env->CallVoidMethod(r, setBenched, (jlong)b.benchedFlags);
        env->CallVoidMethod(r, setResult, (jdouble)b.benchmarkResult);
        env->CallVoidMethod(r, setRatio, (jdouble)b.performanceRatio);
        env->SetObjectArrayElement(out, i, r);
        env->DeleteLocalRef(n); env->DeleteLocalRef(d); env->DeleteLocalRef(im); env->DeleteLocalRef(r);
    }
    return out;
}

NATIVE(jint, createInstance)(JNIEnv* env, jobject, jint tipCount, jint partialsBufferCount, jint compactBufferCount,
                             jint stateCount, jint patternCount, jint eigenBufferCount, jint matrixBufferCount,
                             jint categoryCount, jint scaleBufferCount, jintArray resourceList, jint resourceCount,
                             jlong preferenceFlags, jlong requirementFlags, jobject outDetails) {
    IntIn res(env, resourceList);
    BeagleInstanceDetails det = {0, nullptr, nullptr, nullptr, 0};
    int rc = beagleCreateInstance(tipCount, partialsBufferCount, compactBufferCount, stateCount, patternCount,
                                  eigenBufferCount, matrixBufferCount, categoryCount, scaleBufferCount, res.mut(),
                                  res.p ? resourceCount : 0, (long)preferenceFlags, (long)requirementFlags, &det);
    if (rc >= 0 && outDetails != nullptr) {
        jclass cls = env->FindClass("beagle/InstanceDetails");
        if (cls != nullptr) {
            jmethodID setNum = env->GetMethodID(cls, "setResourceNumber", "(I)V");
            jmethodID setFlags = env->GetMethodID(cls, "setFlags", "(J)V");
            jmethodID setRes = env->GetMethodID(cls, "setResourceName", "(Ljava/lang/String;)V");
            jmethodID setImpl = env->GetMethodID(cls, "setImplementationName", "(Ljava/lang/String;)V");
            if (setNum) env->CallVoidMethod(outDetails, setNum, (jint)det.resourceNumber);
            if (setFlags) env->CallVoidMethod(outDetails, setFlags, (jlong)det.flags);
            if (setRes && det.resourceName) { jstring s = env->NewStringUTF(det.resourceName); env->CallVoidMethod(outDetails, setRes, s); env->DeleteLocalRef(s); }
            if (setImpl && det.implName) { jstring s = env->NewStringUTF(det.implName); env->CallVoidMethod(outDetails, setImpl, s); env->DeleteLocalRef(s); }
        }
    }
    return rc;
}

NATIVE(jint, finalize)(JNIEnv*, jobject, jint instance) { return beagleFinalizeInstance(instance); }
NATIVE(jint, setCPUThreadCount)(JNIEnv*, jobject, jint instance, jint n) { return beagleSetCPUThreadCount(instance, n); }

NATIVE(jint, setPatternWeights)(JNIEnv* env, jobject, jint instance, jdoubleArray w) {
    DblIn a(env, w); return beagleSetPatternWeights(instance, a);
}
NATIVE(jint, setPatternPartitions)(JNIEnv* env, jobject, jint instance, jint partitionCount, jintArray map) {
    IntIn a(env, map); return beagleSetPatternPartitions(instance, partitionCount, a);
}
NATIVE(jint, setTipStates)(JNIEnv* env, jobject, jint instance, jint tip, jintArray states) {
    IntIn a(env, states); return beagleSetTipStates(instance, tip, a);
}
NATIVE(jint, getTipStates)(JNIEnv* env, jobject, jint instance, jint tip, jintArray states) {
    IntOut a(env, states); return beagleGetTipStates(instance, tip, a);
}
NATIVE(jint, setTipPartials)(JNIEnv* env, jobject, jint instance, jint tip, jdoubleArray p) {
    DblIn a(env, p); return beagleSetTipPartials(instance, tip, a);
}
NATIVE(jint, setRootPrePartials)(JNIEnv*, jobject, jint, jintArray, jintArray, jint) { return BEAGLE_ERROR_NO_IMPLEMENTATION; }
NATIVE(jint, setPartials)(JNIEnv* env, jobject, jint instance, jint buffer, jdoubleArray p) {
    DblIn a(env, p); return beagleSetPartials(instance, buffer, a);
}
NATIVE(jint, getPartials)(JNIEnv* env, jobject, jint instance, jint buffer, jint scaleIndex, jdoubleArray out) {
    DblOut a(env, out); return beagleGetPartials(instance, buffer, scaleIndex, a);
}
NATIVE(jint, getLogScaleFactors)(JNIEnv* env, jobject, jint instance, jint scaleIndex, jdoubleArray out) {
    DblOut a(env, out); return beagleGetLogScaleFactors(instance, scaleIndex, a);
}
NATIVE(jint, setEigenDecomposition)(JNIEnv* env, jobject, jint instance, jint idx, jdoubleArray evec, jdoubleArray ievc,
                                    jdoubleArray eval) {
    DblIn a(env, evec), b(env, ievc), c(env, eval);
    return beagleSetEigenDecomposition(instance, idx, a, b, c);
}
NATIVE(jint, setStateFrequencies)(JNIEnv* env, jobject, jint instance, jint idx, jdoubleArray f) {
    DblIn a(env, f); return beagleSetStateFrequencies(instance, idx, a);
}
NATIVE(jint, setCategoryWeights)(JNIEnv* env, jobject, jint instance, jint idx, jdoubleArray w) {
    DblIn a(env, w); return beagleSetCategoryWeights(instance, idx, a);
}
NATIVE(jint, setCategoryRates)(JNIEnv* env, jobject, jint instance, jdoubleArray r) {
    DblIn a(env, r); return beagleSetCategoryRates(instance, a);
}
NATIVE(jint, setCategoryRatesWithIndex)(JNIEnv* env, jobject, jint instance, jint idx, jdoubleArray r) {
    DblIn a(env, r); return beagleSetCategoryRatesWithIndex(instance, idx, a);
}
NATIVE(jint, setTransitionMatrix)(JNIEnv* env, jobject, jint instance, jint idx, jdoubleArray m, jdouble padded) {
    DblIn a(env, m); return beagleSetTransitionMatrix(instance, idx, a, padded);
}
NATIVE(jint, setDifferentialMatrix)(JNIEnv* env, jobject, jint instance, jint idx, jdoubleArray m) {
    DblIn a(env, m); return beagleSetDifferentialMatrix(instance, idx, a);
}
NATIVE(jint, getTransitionMatrix)(JNIEnv* env, jobject, jint instance, jint idx, jdoubleArray m) {
    DblOut a(env, m); return beagleGetTransitionMatrix(instance, idx, a);
}
NATIVE(jint, convolveTransitionMatrices)(JNIEnv* env, jobject, jint instance, jintArray first, jintArray second,
                                         jintArray result, jint count) {
    IntIn a(env, first), b(env, second), c(env, result);
    return beagleConvolveTransitionMatrices(instance, a, b, c, count);
}
NATIVE(jint, addTransitionMatrices)(JNIEnv* env, jobject, jint instance, jintArray first, jintArray second,
                                    jintArray result, jint count) {
    IntIn a(env, first), b(env, second), c(env, result);
    return beagleAddTransitionMatrices(instance, a, b, c, count);
}
NATIVE(jint, transposeTransitionMatrices)(JNIEnv* env, jobject, jint instance, jintArray in, jintArray out, jint count) {
    IntIn a(env, in), b(env, out);
    return beagleTransposeTransitionMatrices(instance, a, b, count);
}
NATIVE(jint, updateTransitionMatrices)(JNIEnv* env, jobject, jint instance, jint eigenIndex, jintArray prob, jintArray d1,
                                       jintArray d2, jdoubleArray lengths, jint count) {
    CritIn<int> a(env, prob), b(env, d1), c(env, d2);
    CritIn<double> t(env, lengths);
    return beagleUpdateTransitionMatrices(instance, eigenIndex, a, b, c, t, count);
}
NATIVE(jint, updateTransitionMatricesWithMultipleModels)(JNIEnv* env, jobject, jint instance, jintArray eigen,
                                                         jintArray rates, jintArray prob, jintArray d1, jintArray d2,
                                                         jdoubleArray lengths, jint count) {
    CritIn<int> e(env, eigen), r(env, rates), a(env, prob), b(env, d1), c(env, d2);
    CritIn<double> t(env, lengths);
    return beagleUpdateTransitionMatricesWithMultipleModels(instance, e, r, a, b, c, t, count);
}
NATIVE(jint, updatePrePartials)(JNIEnv* env, jobject, jint instance, jintArray ops, jint count, jint cum) {
    CritIn<int> a(env, ops);
    return beagleUpdatePrePartials(instance, reinterpret_cast<const BeagleOperation*>((const int*)a), count, cum);
}
NATIVE(jint, updatePrePartialsByPartition)(JNIEnv* env, jobject, jint instance, jintArray ops, jint count) {
    CritIn<int> a(env, ops);
    return beagleUpdatePrePartialsByPartition(instance, reinterpret_cast<const BeagleOperationByPartition*>((const int*)a), count);
}
NATIVE(jint, updatePartials)(JNIEnv* env, jobject, jint instance, jintArray ops, jint count, jint cum) {
    CritIn<int> a(env, ops);
    return beagleUpdatePartials(instance, reinterpret_cast<const BeagleOperation*>((const int*)a), count, cum);
}
NATIVE(jint, updatePartialsByPartition)(JNIEnv* env, jobject, jint instance, jintArray ops, jint count) {
    CritIn<int> a(env, ops);
    return beagleUpdatePartialsByPartition(instance, reinterpret_cast<const BeagleOperationByPartition*>((const int*)a), count);
}
NATIVE(jint, waitForPartials)(JNIEnv* env, jobject, jint instance, jintArray dest, jint count) {
    IntIn a(env, dest); return beagleWaitForPartials(instance, a, count);
}
NATIVE(jint, accumulateScaleFactors)(JNIEnv* env, jobject, jint instance, jintArray idx, jint count, jint cum) {
    CritIn<int> a(env, idx); return beagleAccumulateScaleFactors(instance, a, count, cum);
}
NATIVE(jint, accumulateScaleFactorsByPartition)(JNIEnv* env, jobject, jint instance, jintArray idx, jint count, jint cum,
                                                jint part) {
    CritIn<int> a(env, idx); return beagleAccumulateScaleFactorsByPartition(instance, a, count, cum, part);
}
NATIVE(jint, removeScaleFactors)(JNIEnv* env, jobject, jint instance, jintArray idx, jint count, jint cum) {
    IntIn a(env, idx); return beagleRemoveScaleFactors(instance, a, count, cum);
}
NATIVE(jint, removeScaleFactorsByPartition)(JNIEnv* env, jobject, jint instance, jintArray idx, jint count, jint cum,
                                            jint part) {
    IntIn a(env, idx); return beagleRemoveScaleFactorsByPartition(instance, a, count, cum, part);
}
NATIVE(jint, resetScaleFactors)(JNIEnv*, jobject, jint instance, jint cum) { return beagleResetScaleFactors(instance, cum); }
NATIVE(jint, resetScaleFactorsByPartition)(JNIEnv*, jobject, jint instance, jint cum, jint part) {
    return beagleResetScaleFactorsByPartition(instance, cum, part);
}
NATIVE(jint, copyScaleFactors)(JNIEnv*, jobject, jint instance, jint dest, jint src) {
    return beagleCopyScaleFactors(instance, dest, src);
}
NATIVE(jint, calculateRootLogLikelihoods)(JNIEnv* env, jobject, jint instance, jintArray buffers, jintArray weights,
                                          jintArray freqs, jintArray scale, jint count, jdoubleArray out) {
    IntIn a(env, buffers), b(env, weights), c(env, freqs), d(env, scale);
    DblOut o(env, out);
    return beagleCalculateRootLogLikelihoods(instance, a, b, c, d, count, o);
}
NATIVE(jint, calculateRootLogLikelihoodsByPartition)(JNIEnv* env, jobject, jint instance, jintArray buffers,
                                                     jintArray weights, jintArray freqs, jintArray scale,
                                                     jintArray parts, jint partitionCount, jint count,
                                                     jdoubleArray outByPartition, jdoubleArray out) {
    IntIn a(env, buffers), b(env, weights), c(env, freqs), d(env, scale), p(env, parts);
    DblOut o1(env, outByPartition), o2(env, out);
    return beagleCalculateRootLogLikelihoodsByPartition(instance, a, b, c, d, p, partitionCount, count, o1, o2);
}
NATIVE(jint, getSiteLogLikelihoods)(JNIEnv* env, jobject, jint instance, jdoubleArray out) {
    DblOut o(env, out); return beagleGetSiteLogLikelihoods(instance, o);
}
// remaining derivative API: SURVEY.md 8f "next" rows
NATIVE(jint, calculateEdgeDifferentials)(JNIEnv* env, jobject, jint instance, jintArray post, jintArray pre, jintArray deriv,
                                         jintArray weights, jint count, jdoubleArray out, jdoubleArray outSum,
                                         jdoubleArray outSumSquared) {
    IntIn a(env, post), b(env, pre), c(env, deriv), w(env, weights);
    DblOut o(env, out), s1(env, outSum), s2(env, outSumSquared);
    return beagleCalculateEdgeDerivatives(instance, a, b, c, w, count, o, s1, s2);
}
NATIVE(jint, calculateCrossProductDifferentials)(JNIEnv* env, jobject, jint instance, jintArray post, jintArray pre,
                                                 jintArray rates, jintArray weights, jdoubleArray lengths, jint count,
                                                 jdoubleArray outSum, jdoubleArray outSumSquared) {
    IntIn a(env, post), b(env, pre), r(env, rates), w(env, weights);
    DblIn t(env, lengths);
    DblOut s1(env, outSum), s2(env, outSumSquared);
    return beagleCalculateCrossProductDerivative(instance, a, b, r, w, t, count, s1, s2);
}
NATIVE(jint, calculateEdgeDerivative)(JNIEnv*, jobject, jint, jintArray, jintArray, jint, jintArray, jintArray, jint, jint,
                                      jint, jintArray, jint, jdoubleArray, jdoubleArray) {
    return BEAGLE_ERROR_NO_IMPLEMENTATION;
}
