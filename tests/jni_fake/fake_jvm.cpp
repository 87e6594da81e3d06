// fake_jvm.cpp -- test harness: a fake JNIEnv (function table over plain C++ objects) that drives the
// exported Java_beagle_BeagleJNIWrapper_* symbols of libhmsbeagle-jni.so without a JVM
// (SURVEY.md Appendix C).  Modes:
//   fake_jvm info                  -> getVersion / getResourceList through JNI (no GPU needed)
//   fake_jvm tiny <fixture.txt>    -> the BEAGLE "tiny test" call sequence (BeagleFactory.main of the jar):
//                                     prints logL (needs a GPU)
#include "jni_min.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <string>
#include <vector>

struct FakeObj {
    std::string kind, cls, str;
    std::vector<jint> ints;
    std::vector<jdouble> dbls;
    std::vector<FakeObj*> elems;
    std::map<std::string, std::string> sfields;
    std::map<std::string, long long> ifields;
    std::map<std::string, double> dfields;
};
struct FakeMethod { std::string name, sig; };

static FakeObj* O(jobject o) { return reinterpret_cast<FakeObj*>(o); }
static jobject J(FakeObj* o) { return reinterpret_cast<jobject>(o); }

static jint f_GetVersion(JNIEnv*) { return JNI_VERSION_1_6; }
static jclass f_FindClass(JNIEnv*, const char* name) { auto* c = new FakeObj; c->kind = "class"; c->cls = name; return J(c); }
static void f_ExceptionClear(JNIEnv*) {}
static jboolean f_ExceptionCheck(JNIEnv*) { return JNI_FALSE; }
static void f_DeleteLocalRef(JNIEnv*, jobject) {}
static jmethodID f_GetMethodID(JNIEnv*, jclass, const char* name, const char* sig) {
    return reinterpret_cast<jmethodID>(new FakeMethod{name, sig});
}
static void applyCall(FakeObj* obj, FakeMethod* m, va_list ap) {
    if (m->sig == "(I)V") obj->ifields[m->name] = va_arg(ap, jint);
    else if (m->sig == "(J)V") obj->ifields[m->name] = va_arg(ap, jlong);
    else if (m->sig == "(D)V") obj->dfields[m->name] = va_arg(ap, jdouble);
    else if (m->sig == "(Ljava/lang/String;)V") { FakeObj* s = O(va_arg(ap, jobject)); obj->sfields[m->name] = s ? s->str : ""; }
    else { fprintf(stderr, "fake_jvm: unsupported signature %s\n", m->sig.c_str()); exit(2); }
}
static jobject f_NewObject(JNIEnv*, jclass c, jmethodID mid, ...) {
    auto* o = new FakeObj; o->kind = "object"; o->cls = O(c)->cls;
    va_list ap; va_start(ap, mid); applyCall(o, reinterpret_cast<FakeMethod*>(mid), ap); va_end(ap);
    return J(o);
}
static void f_CallVoidMethod(JNIEnv*, jobject obj, jmethodID mid, ...) {
    va_list ap; va_start(ap, mid); applyCall(O(obj), reinterpret_cast<FakeMethod*>(mid), ap); va_end(ap);
}
static jstring f_NewStringUTF(JNIEnv*, const char* s) { auto* o = new FakeObj; o->kind = "string"; o->str = s ? s : ""; return J(o); }
static jsize f_GetArrayLength(JNIEnv*, jarray a) {
    FakeObj* o = O(a);
    return (jsize)(o->kind == "int[]" ? o->ints.size() : o->kind == "double[]" ? o->dbls.size() : o->elems.size());
}
static jobjectArray f_NewObjectArray(JNIEnv*, jsize n, jclass, jobject) { auto* o = new FakeObj; o->kind = "object[]"; o->elems.assign(n, nullptr); return J(o); }
static void f_SetObjectArrayElement(JNIEnv*, jobjectArray a, jsize i, jobject v) { O(a)->elems[i] = O(v); }
static int g_critical = 0;      // open GetPrimitiveArrayCritical regions
// element access hands out a COPY and honours the release mode, like a JVM that does not pin
static jint* f_GetIntArrayElements(JNIEnv*, jintArray a, jboolean* isCopy) {
    if (g_critical) { fprintf(stderr, "JNI call inside a critical region\n"); abort(); }
    if (isCopy) *isCopy = JNI_TRUE;
    auto& v = O(a)->ints; jint* c = (jint*)malloc(sizeof(jint) * (v.size() + 1)); memcpy(c, v.data(), sizeof(jint) * v.size()); return c;
}
static void f_ReleaseIntArrayElements(JNIEnv*, jintArray a, jint* e, jint mode) {
    auto& v = O(a)->ints;
    if (mode != JNI_ABORT) memcpy(v.data(), e, sizeof(jint) * v.size());
    if (mode != JNI_COMMIT) free(e);
}
static jdouble* f_GetDoubleArrayElements(JNIEnv*, jdoubleArray a, jboolean* isCopy) {
    if (isCopy) *isCopy = JNI_TRUE;
    auto& v = O(a)->dbls; jdouble* c = (jdouble*)malloc(sizeof(jdouble) * (v.size() + 1)); memcpy(c, v.data(), sizeof(jdouble) * v.size()); return c;
}
// critical access pins: the array's own storage is handed out; a counter checks that every Get is released and that no
// other JNI call happens in between (the specification's rule)
static void* f_GetPrimitiveArrayCritical(JNIEnv*, jarray a, jboolean* isCopy) {
    if (isCopy) *isCopy = JNI_FALSE;
    ++g_critical;
    FakeObj* o = O(a);
    return o->kind == "int[]" ? (void*)o->ints.data() : (void*)o->dbls.data();
}
static void f_ReleasePrimitiveArrayCritical(JNIEnv*, jarray, void*, jint) {
    if (--g_critical < 0) { fprintf(stderr, "unbalanced ReleasePrimitiveArrayCritical\n"); abort(); }
}
static void f_ReleaseDoubleArrayElements(JNIEnv*, jdoubleArray a, jdouble* e, jint mode) {
    auto& v = O(a)->dbls;
    if (mode != JNI_ABORT) memcpy(v.data(), e, sizeof(jdouble) * v.size());
    if (mode != JNI_COMMIT) free(e);
}

static jintArray IA(std::vector<jint> v) { auto* o = new FakeObj; o->kind = "int[]"; o->ints = std::move(v); return J(o); }
static jdoubleArray DA(std::vector<jdouble> v) { auto* o = new FakeObj; o->kind = "double[]"; o->dbls = std::move(v); return J(o); }

template <typename F> static F sym(void* h, const char* name) {
    std::string full = std::string("Java_beagle_BeagleJNIWrapper_") + name;
    void* p = dlsym(h, full.c_str());
    if (!p) { fprintf(stderr, "missing symbol %s\n", full.c_str()); exit(3); }
    return reinterpret_cast<F>(p);
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: fake_jvm <libhmsbeagle-jni.so> info|tiny|gradient [fixture]\n"); return 1; }
    void* h = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
    static JNINativeInterface_ table;
    memset(&table, 0, sizeof table);
    table.slot[JNI_IDX_GetVersion] = (void*)f_GetVersion;
    table.slot[JNI_IDX_FindClass] = (void*)f_FindClass;
    table.slot[JNI_IDX_ExceptionClear] = (void*)f_ExceptionClear;
    table.slot[JNI_IDX_ExceptionCheck] = (void*)f_ExceptionCheck;
    table.slot[JNI_IDX_DeleteLocalRef] = (void*)f_DeleteLocalRef;
    table.slot[JNI_IDX_GetMethodID] = (void*)f_GetMethodID;
    table.slot[JNI_IDX_NewObject] = (void*)f_NewObject;
    table.slot[JNI_IDX_CallVoidMethod] = (void*)f_CallVoidMethod;
    table.slot[JNI_IDX_NewStringUTF] = (void*)f_NewStringUTF;
    table.slot[JNI_IDX_GetArrayLength] = (void*)f_GetArrayLength;
    table.slot[JNI_IDX_NewObjectArray] = (void*)f_NewObjectArray;
    table.slot[JNI_IDX_SetObjectArrayElement] = (void*)f_SetObjectArrayElement;
    table.slot[JNI_IDX_GetPrimitiveArrayCritical] = (void*)f_GetPrimitiveArrayCritical;
    table.slot[JNI_IDX_ReleasePrimitiveArrayCritical] = (void*)f_ReleasePrimitiveArrayCritical;
    table.slot[JNI_IDX_GetIntArrayElements] = (void*)f_GetIntArrayElements;
    table.slot[JNI_IDX_ReleaseIntArrayElements] = (void*)f_ReleaseIntArrayElements;
    table.slot[JNI_IDX_GetDoubleArrayElements] = (void*)f_GetDoubleArrayElements;
    table.slot[JNI_IDX_ReleaseDoubleArrayElements] = (void*)f_ReleaseDoubleArrayElements;
    JNIEnv envObj; envObj.functions = &table;
    JNIEnv* env = &envObj;
    jobject self = J(new FakeObj);

    auto getVersion = sym<jstring (*)(JNIEnv*, jobject)>(h, "getVersion");
    auto getResourceList = sym<jobjectArray (*)(JNIEnv*, jobject)>(h, "getResourceList");
    printf("version %s\n", O(getVersion(env, self))->str.c_str());
    FakeObj* rl = O(getResourceList(env, self));
    printf("resources %zu\n", rl->elems.size());
    for (FakeObj* r : rl->elems)
        printf("resource %lld name=%s flags=%lld\n", r->ifields["<init>"], r->sfields["setName"].c_str(), r->ifields["setFlags"]);
    if (std::string(argv[2]) == "info") return 0;
    if (std::string(argv[2]) == "auto") {
        // BDLD:413-426 (-beagle_auto): the benchmarked list, fastest first
        typedef jobjectArray (*bench_t)(JNIEnv*, jobject, jint, jint, jint, jint, jint, jintArray, jint, jlong, jlong, jint, jint, jint, jlong);
        jobjectArray arr = sym<bench_t>(h, "getBenchmarkedResourceList")(env, self, 16, 16, 4, 500, 4, nullptr, 0, 0, 0, 1, 1, 0, 1);
        if (arr == nullptr) { printf("benchmarked null\n"); return 0; }
        for (FakeObj* r : O(arr)->elems)
            printf("benchmarked %lld resource=%lld name=%s impl=%s rc=%lld ms=%.4f ratio=%.3f\n", r->ifields["<init>"],
                   r->ifields["setResourceNumber"], r->sfields["setName"].c_str(), r->sfields["setImplName"].c_str(),
                   r->ifields["setReturnCode"], r->dfields["setBenchmarkResult"], r->dfields["setPerformanceRatio"]);
        return 0;
    }

    // ---- tiny test -----------------------------------------------------------------------------
    FILE* f = fopen(argv[3], "r");
    int nTips, nPat;
    if (!f || fscanf(f, "%d %d", &nTips, &nPat) != 2 || nTips != 3) { fprintf(stderr, "bad fixture\n"); return 1; }
    std::vector<std::vector<jint>> states(nTips, std::vector<jint>(nPat));
    for (auto& s : states) for (auto& v : s) if (fscanf(f, "%d", &v) != 1) return 1;
    fclose(f);
    typedef jint (*create_t)(JNIEnv*, jobject, jint, jint, jint, jint, jint, jint, jint, jint, jint, jintArray, jint, jlong, jlong, jobject);
    jobject details = J(new FakeObj);
    // 2 post-order + 5 pre-order partials buffers (indices 3,4 and 5..9), 4 branch matrices + 1 differential matrix
    jint inst = sym<create_t>(h, "createInstance")(env, self, 3, 7, 3, 4, nPat, 1, 5, 1, 0, IA({1, 0}), 2, 0, 0, details);
    if (inst < 0) { fprintf(stderr, "createInstance returned %d\n", inst); return 4; }
    printf("instance %d resource=%lld impl=%s flags=%lld\n", inst, O(details)->ifields["setResourceNumber"],
           O(details)->sfields["setImplementationName"].c_str(), O(details)->ifields["setFlags"]);
    auto setTipStates = sym<jint (*)(JNIEnv*, jobject, jint, jint, jintArray)>(h, "setTipStates");
    for (int t = 0; t < 3; ++t) if (setTipStates(env, self, inst, t, IA(states[t]))) return 5;
    if (sym<jint (*)(JNIEnv*, jobject, jint, jdoubleArray)>(h, "setPatternWeights")(env, self, inst, DA(std::vector<jdouble>(nPat, 1.0)))) return 5;
    // JC69 eigen system: the constants of beagle.BeagleFactory.main
    std::vector<jdouble> evec = {1.0, 2.0, 0.0, 0.5, 1.0, -2.0, 0.5, 0.0, 1.0, 2.0, 0.0, -0.5, 1.0, -2.0, -0.5, 0.0};
    std::vector<jdouble> ivec = {0.25, 0.25, 0.25, 0.25, 0.125, -0.125, 0.125, -0.125, 0.0, 1.0, 0.0, -1.0, 1.0, 0.0, -1.0, 0.0};
    std::vector<jdouble> eval = {0.0, -1.3333333333333333, -1.3333333333333333, -1.3333333333333333};
    if (sym<jint (*)(JNIEnv*, jobject, jint, jint, jdoubleArray, jdoubleArray, jdoubleArray)>(h, "setEigenDecomposition")(env, self, inst, 0, DA(evec), DA(ivec), DA(eval))) return 5;
    if (sym<jint (*)(JNIEnv*, jobject, jint, jdoubleArray)>(h, "setCategoryRates")(env, self, inst, DA({1.0}))) return 5;
    if (sym<jint (*)(JNIEnv*, jobject, jint, jint, jdoubleArray)>(h, "setCategoryWeights")(env, self, inst, 0, DA({1.0}))) return 5;
    if (sym<jint (*)(JNIEnv*, jobject, jint, jint, jdoubleArray)>(h, "setStateFrequencies")(env, self, inst, 0, DA({0.25, 0.25, 0.25, 0.25}))) return 5;
    // derivative index arrays are Java null on the hot path
    if (sym<jint (*)(JNIEnv*, jobject, jint, jint, jintArray, jintArray, jintArray, jdoubleArray, jint)>(h, "updateTransitionMatrices")(
            env, self, inst, 0, IA({0, 1, 2, 3}), nullptr, nullptr, DA({0.1, 0.1, 0.2, 0.1}), 4)) return 6;
    // ops: node 3 = (tip0, tip1), root 4 = (tip2, node3)   [dest, sw, sr, c1, m1, c2, m2]
    if (sym<jint (*)(JNIEnv*, jobject, jint, jintArray, jint, jint)>(h, "updatePartials")(
            env, self, inst, IA({3, -1, -1, 0, 0, 1, 1, 4, -1, -1, 2, 2, 3, 3}), 2, -1)) return 7;
    jdoubleArray out = DA({0.0});
    jint rc = sym<jint (*)(JNIEnv*, jobject, jint, jintArray, jintArray, jintArray, jintArray, jint, jdoubleArray)>(h, "calculateRootLogLikelihoods")(
        env, self, inst, IA({4}), IA({0}), IA({0}), IA({-1}), 1, out);
    printf("logL %.5f rc %d\n", O(out)->dbls[0], rc);
    jdoubleArray site = DA(std::vector<jdouble>(nPat, 0.0));
    sym<jint (*)(JNIEnv*, jobject, jint, jdoubleArray)>(h, "getSiteLogLikelihoods")(env, self, inst, site);
    double s = 0; for (double v : O(site)->dbls) s += v;
    printf("siteSum %.5f\n", s);
    if (std::string(argv[2]) == "gradient") {
        // the pre-order / derivative natives (AbstractBeagleGradientDelegate, SubstitutionModelCrossProductDelegate):
        // pre-order buffer of node k = 5 + k, root = node 4
        std::vector<jdouble> rootPre(4 * nPat, 0.25);
        if (sym<jint (*)(JNIEnv*, jobject, jint, jint, jdoubleArray)>(h, "setPartials")(env, self, inst, 9, DA(rootPre))) return 8;
        // [dest pre, sw, sr, parent pre, own matrix, sibling post, sibling matrix]
        if (sym<jint (*)(JNIEnv*, jobject, jint, jintArray, jint, jint)>(h, "updatePrePartials")(
                env, self, inst, IA({8, -1, -1, 9, 3, 2, 2,   7, -1, -1, 9, 2, 3, 3,   5, -1, -1, 8, 0, 1, 1,   6, -1, -1, 8, 1, 0, 0}), 4, -1)) return 8;
        std::vector<jdouble> Q(16, 1.0 / 3.0);
        for (int i = 0; i < 4; ++i) Q[5 * i] = -1.0;
        if (sym<jint (*)(JNIEnv*, jobject, jint, jint, jdoubleArray)>(h, "setDifferentialMatrix")(env, self, inst, 4, DA(Q))) return 8;
        jdoubleArray sum = DA(std::vector<jdouble>(4, 0.0)), sumSq = DA(std::vector<jdouble>(4, 0.0));
        rc = sym<jint (*)(JNIEnv*, jobject, jint, jintArray, jintArray, jintArray, jintArray, jint, jdoubleArray, jdoubleArray, jdoubleArray)>(
            h, "calculateEdgeDifferentials")(env, self, inst, IA({0, 1, 2, 3}), IA({5, 6, 7, 8}), IA({4, 4, 4, 4}), IA({0}), 4,
                                             nullptr, sum, sumSq);
        printf("edge rc %d %.9f %.9f %.9f %.9f\n", rc, O(sum)->dbls[0], O(sum)->dbls[1], O(sum)->dbls[2], O(sum)->dbls[3]);
        jdoubleArray cross = DA(std::vector<jdouble>(16, 1.0));      // the call ADDS to what it is given
        rc = sym<jint (*)(JNIEnv*, jobject, jint, jintArray, jintArray, jintArray, jintArray, jdoubleArray, jint, jdoubleArray, jdoubleArray)>(
            h, "calculateCrossProductDifferentials")(env, self, inst, IA({0, 1, 2, 3}), IA({5, 6, 7, 8}), IA({0}), IA({0}),
                                                     DA({0.1, 0.1, 0.2, 0.1}), 4, cross, nullptr);
        printf("cross rc %d", rc);
        for (double v : O(cross)->dbls) printf(" %.9f", v - 1.0);
        printf("\n");
    }
    sym<jint (*)(JNIEnv*, jobject, jint)>(h, "finalize")(env, self, inst);
    return 0;
}
