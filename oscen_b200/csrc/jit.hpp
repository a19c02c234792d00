/*
 * jit.hpp — run-time compilation of generated fused kernels (codegen.hpp): NVRTC (dlopen'd
 * libnvrtc.so.12) -> sm_100a cubin -> driver-API module (dlopen'd libcuda.so.1) -> CUfunction
 * launched on the engine's stream.  Neither library is a link-time dependency: an engine that only
 * runs the built-in kernels never loads them, and a host without them gets a clean error from
 * create(), never a different code path.
 *
 * Compile flags mirror the in-tree build (oscen_b200/build.py): -arch=sm_100a, --fmad=false (rustc
 * never contracts a*b+c), IEEE division / sqrt, no fast-math, -lineinfo.  The generated translation
 * unit includes the engine's own device headers, so the node bodies are the very same functions the
 * hand-written kernels call.  The build embeds those headers in the library as text (jit_headers.inc)
 * and hands them to nvrtcCreateProgram: a relocated .so needs no source tree.  $OSCEN_B200_CSRC
 * overrides the embedded copies with a directory (development: try a header edit without a rebuild).
 */
#pragma once
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace oscen {
namespace jit {

/* the device headers a generated kernel includes, embedded as text by the build (oscen_b200/build.py writes
   jit_headers.inc): run-time compilation then needs no source tree next to the shared library */
struct EmbeddedHeader { const char *name; const char *text; };
}  // namespace jit
}  // namespace oscen
#if defined(__has_include)
#if __has_include("jit_headers.inc")
#include "jit_headers.inc"
#define OSCEN_B200_HAVE_EMBEDDED_HEADERS 1
#endif
#endif
namespace oscen {
namespace jit {

/* ---- NVRTC (subset) */
struct Nvrtc {
    void *lib = nullptr;
    int (*CreateProgram)(void **, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    int (*DestroyProgram)(void **) = nullptr;
    int (*CompileProgram)(void *, int, const char *const *) = nullptr;
    int (*GetCUBINSize)(void *, size_t *) = nullptr;
    int (*GetCUBIN)(void *, char *) = nullptr;
    int (*GetProgramLogSize)(void *, size_t *) = nullptr;
    int (*GetProgramLog)(void *, char *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
inline Nvrtc &nvrtc() { static Nvrtc n; return n; }

inline bool nvrtc_load(std::string &err) {
    Nvrtc &n = nvrtc();
    if (n.lib) return true;
    const char *names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char *nm : names) { n.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL); if (n.lib) break; }
    if (!n.lib) { err = "NVRTC (libnvrtc.so.12) cannot be loaded: generated fused kernels need it; PLAN_NODES does not"; return false; }
#define OSC_SYM(field, name) *(void **)(&n.field) = dlsym(n.lib, name)
    OSC_SYM(CreateProgram, "nvrtcCreateProgram"); OSC_SYM(DestroyProgram, "nvrtcDestroyProgram");
    OSC_SYM(CompileProgram, "nvrtcCompileProgram"); OSC_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
    OSC_SYM(GetCUBIN, "nvrtcGetCUBIN"); OSC_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    OSC_SYM(GetProgramLog, "nvrtcGetProgramLog"); OSC_SYM(GetErrorString, "nvrtcGetErrorString");
#undef OSC_SYM
    if (!n.CreateProgram || !n.CompileProgram || !n.GetCUBINSize || !n.GetCUBIN) { err = "NVRTC symbols missing"; dlclose(n.lib); n.lib = nullptr; return false; }
    return true;
}

/* directory of the engine's device headers: $OSCEN_B200_CSRC, else <dir of this .so>/../csrc */
inline std::string csrc_dir() {
    if (const char *e = getenv("OSCEN_B200_CSRC")) return e;
    Dl_info info;
    if (dladdr((void *)&csrc_dir, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        const size_t k = p.rfind('/');
        p = k == std::string::npos ? std::string(".") : p.substr(0, k);
        return p + "/../csrc";
    }
    return "oscen_b200/csrc";
}

/* source -> sm_100a cubin.  Needs no GPU (the CPU test suite compiles every test graph).  Compiled images are kept
   per process, keyed by the source text: engines of the same graph share one compilation. */
inline bool compile_cubin_uncached(const std::string &source, std::vector<char> &cubin, std::string &log, std::string &err);
inline bool compile_cubin(const std::string &source, std::vector<char> &cubin, std::string &log, std::string &err) {
    static std::mutex mu;
    static std::map<std::string, std::vector<char>> cache;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(source);
        if (it != cache.end()) { cubin = it->second; return true; }
    }
    if (!compile_cubin_uncached(source, cubin, log, err)) return false;
    std::lock_guard<std::mutex> lock(mu);
    cache[source] = cubin;
    return true;
}
inline bool compile_cubin_uncached(const std::string &source, std::vector<char> &cubin, std::string &log, std::string &err) {
    if (!nvrtc_load(err)) return false;
    Nvrtc &n = nvrtc();
    void *prog = nullptr;
    /* headers: the embedded copies (the ones this library was built with), unless $OSCEN_B200_CSRC points at a source tree */
    std::vector<const char *> h_text, h_name;
#ifdef OSCEN_B200_HAVE_EMBEDDED_HEADERS
    if (!getenv("OSCEN_B200_CSRC"))
        for (int i = 0; i < kNJitHeaders; ++i) { h_text.push_back(kJitHeaders[i].text); h_name.push_back(kJitHeaders[i].name); }
#endif
    int rc = n.CreateProgram(&prog, source.c_str(), "oscen_b200_generated.cu", (int)h_text.size(), h_text.empty() ? nullptr : h_text.data(),
                             h_name.empty() ? nullptr : h_name.data());
    if (rc != 0) { err = std::string("nvrtcCreateProgram: ") + (n.GetErrorString ? n.GetErrorString(rc) : "error"); return false; }
    const std::string inc = "--include-path=" + (h_text.empty() ? csrc_dir() : std::string("/nonexistent-oscen-b200-headers-are-embedded"));
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "--prec-div=true", "--prec-sqrt=true",
                          "-lineinfo", "-default-device", inc.c_str()};
    rc = n.CompileProgram(prog, (int)(sizeof opts / sizeof opts[0]), opts);
    size_t ls = 0;
    if (n.GetProgramLogSize && n.GetProgramLogSize(prog, &ls) == 0 && ls > 1) { log.resize(ls); n.GetProgramLog(prog, &log[0]); }
    if (rc != 0) {
        err = std::string("nvrtcCompileProgram: ") + (n.GetErrorString ? n.GetErrorString(rc) : "error") + "\n" + log;
        n.DestroyProgram(&prog);
        return false;
    }
    size_t cs = 0;
    rc = n.GetCUBINSize(prog, &cs);
    if (rc == 0 && cs) { cubin.resize(cs); rc = n.GetCUBIN(prog, cubin.data()); }
    n.DestroyProgram(&prog);
    if (rc != 0 || cubin.empty()) { err = "nvrtcGetCUBIN failed"; return false; }
    return true;
}

/* ---- driver API (subset) */
struct Driver {
    void *lib = nullptr;
    int (*ModuleLoadData)(void **, const void *) = nullptr;
    int (*ModuleUnload)(void *) = nullptr;
    int (*ModuleGetFunction)(void **, void *, const char *) = nullptr;
    int (*LaunchKernel)(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void *, void **, void **) = nullptr;
    int (*GetErrorString)(int, const char **) = nullptr;
    int (*FuncSetAttribute)(void *, int, int) = nullptr;
};
inline Driver &driver() { static Driver d; return d; }
inline bool driver_load(std::string &err) {
    Driver &d = driver();
    if (d.lib) return true;
    d.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!d.lib) { err = "libcuda.so.1 cannot be loaded"; return false; }
#define OSC_SYM(field, name) *(void **)(&d.field) = dlsym(d.lib, name)
    OSC_SYM(ModuleLoadData, "cuModuleLoadData"); OSC_SYM(ModuleUnload, "cuModuleUnload");
    OSC_SYM(ModuleGetFunction, "cuModuleGetFunction"); OSC_SYM(LaunchKernel, "cuLaunchKernel");
    OSC_SYM(GetErrorString, "cuGetErrorString"); OSC_SYM(FuncSetAttribute, "cuFuncSetAttribute");
#undef OSC_SYM
    if (!d.ModuleLoadData || !d.ModuleGetFunction || !d.LaunchKernel) { err = "driver API symbols missing"; return false; }
    return true;
}
inline std::string driver_error(int rc) {
    const char *s = nullptr;
    if (driver().GetErrorString && driver().GetErrorString(rc, &s) == 0 && s) return s;
    return "CUDA driver error " + std::to_string(rc);
}

struct Kernel {
    void *module = nullptr, *function = nullptr;
    std::vector<char> cubin;
    std::string source, log;
    unsigned dyn_smem = 0;               /* dynamic shared memory per CTA (FIR rings of [sinc] cross-rate edges) */
    /* the runtime's primary context must be current on this thread (cudaSetDevice + any runtime call) */
    bool load(const char *name, std::string &err) {
        if (!driver_load(err)) return false;
        int rc = driver().ModuleLoadData(&module, cubin.data());
        if (rc != 0) { err = "cuModuleLoadData: " + driver_error(rc); return false; }
        rc = driver().ModuleGetFunction(&function, module, name);
        if (rc != 0) { err = "cuModuleGetFunction: " + driver_error(rc); return false; }
        if (dyn_smem > 0u) {               /* CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES = 8 (static tiles + rings may pass 48 KB) */
            if (!driver().FuncSetAttribute) { err = "cuFuncSetAttribute missing"; return false; }
            rc = driver().FuncSetAttribute(function, 8, (int)dyn_smem);
            if (rc != 0) { err = "cuFuncSetAttribute(max dynamic shared memory " + std::to_string(dyn_smem) + " B): " + driver_error(rc); return false; }
        }
        return true;
    }
    int launch(unsigned grid, unsigned block, void *stream, void *kernel_args_struct) const {
        void *params[] = {kernel_args_struct};
        return driver().LaunchKernel(function, grid, 1, 1, block, 1, 1, dyn_smem, stream, params, nullptr);
    }
    void unload() { if (module && driver().ModuleUnload) driver().ModuleUnload(module); module = nullptr; function = nullptr; }
};

}  // namespace jit
}  // namespace oscen
