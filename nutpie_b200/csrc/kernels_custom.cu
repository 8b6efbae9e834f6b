// kernels_custom.cu — NB200_MODEL_CUSTOM: the sampler kernel compiled at run time around a
// user-supplied device density.
//
// The reference turns a model into machine code when the model is compiled (numba cfunc,
// python/nutpie/compile_pymc.py:970-1006; BridgeStan shared object, compile_stan.py:17-130) and
// hands nuts-rs a function POINTER (src/pymc.rs:50-62).  A device engine cannot call through a
// host pointer per leapfrog, and an indirect device call per gradient would forbid inlining the
// density into the integrator.  So the equivalent step here is: NVRTC compiles
//     [amalgamated engine headers] + [user source defining nb200_user_logp]
// into a cubin for sm_100a whose only kernels are nuts_kernel<CustomModel, W, NIT> and
// component_kernel<CustomModel, W> for the geometry the host picked; the cubin is loaded with the
// runtime's library API and launched exactly like the statically compiled kernels.
//
// libnvrtc is dlopen'ed on first use so that libnutpie_b200.so itself has no load-time
// dependency on it (the built-in densities never need it).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "launch_impl.cuh"

namespace nb200 {

// the engine's device headers, inlined into one string by nutpie_b200/build.py
static const char* const kAmalgam =
#include "../build/rtc_amalgam.inc"
    ;

namespace {

struct Nvrtc {
    void* h = nullptr;
    std::string err;
    decltype(&nvrtcCreateProgram) createProgram = nullptr;
    decltype(&nvrtcDestroyProgram) destroyProgram = nullptr;
    decltype(&nvrtcCompileProgram) compileProgram = nullptr;
    decltype(&nvrtcGetProgramLogSize) getProgramLogSize = nullptr;
    decltype(&nvrtcGetProgramLog) getProgramLog = nullptr;
    decltype(&nvrtcGetCUBINSize) getCUBINSize = nullptr;
    decltype(&nvrtcGetCUBIN) getCUBIN = nullptr;
    decltype(&nvrtcAddNameExpression) addNameExpression = nullptr;
    decltype(&nvrtcGetLoweredName) getLoweredName = nullptr;
    decltype(&nvrtcGetErrorString) getErrorString = nullptr;
};

Nvrtc& nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        std::vector<std::string> cands;
        if (const char* e = std::getenv("NB200_NVRTC_PATH")) cands.push_back(e);
        cands.push_back("libnvrtc.so.12");
        cands.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
        cands.push_back("libnvrtc.so");
        for (const auto& c : cands) {
            n.h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
            if (n.h) break;
        }
        if (!n.h) {
            n.err = "libnvrtc.so.12 not found (set NB200_NVRTC_PATH): custom CUDA densities need NVRTC";
            return;
        }
#define NB_SYM(field, name)                                                    \
    n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.h, name));           \
    if (!n.field) n.err = std::string("libnvrtc lacks ") + name;
        NB_SYM(createProgram, "nvrtcCreateProgram")
        NB_SYM(destroyProgram, "nvrtcDestroyProgram")
        NB_SYM(compileProgram, "nvrtcCompileProgram")
        NB_SYM(getProgramLogSize, "nvrtcGetProgramLogSize")
        NB_SYM(getProgramLog, "nvrtcGetProgramLog")
        NB_SYM(getCUBINSize, "nvrtcGetCUBINSize")
        NB_SYM(getCUBIN, "nvrtcGetCUBIN")
        NB_SYM(addNameExpression, "nvrtcAddNameExpression")
        NB_SYM(getLoweredName, "nvrtcGetLoweredName")
        NB_SYM(getErrorString, "nvrtcGetErrorString")
#undef NB_SYM
    });
    return n;
}

struct Built {  // one (W, NIT) specialisation of one source
    std::vector<char> cubin;
    std::string nuts_name, comp_name;
    cudaLibrary_t lib = nullptr;  // loaded lazily: compiling needs no GPU, loading does
    cudaKernel_t nuts = nullptr, comp = nullptr;
};

struct Program {
    std::string source;
    std::map<std::pair<int, int>, std::unique_ptr<Built>> built;
};

std::mutex g_mu;
std::vector<std::unique_ptr<Program>> g_programs;
thread_local std::string t_log;

int compile_locked(Program& prog, int W, int NIT, Built** out) {
    auto key = std::make_pair(W, NIT);
    auto it = prog.built.find(key);
    if (it != prog.built.end()) {
        *out = it->second.get();
        return 0;
    }
    Nvrtc& n = nvrtc();
    if (!n.err.empty()) {
        t_log = n.err;
        return -1;
    }
    std::string src;
    src.reserve(prog.source.size() + 200000);
    src += kAmalgam;
    src += "\n#line 1 \"nb200_user_density.cu\"\n";
    src += prog.source;
    src += "\n";
    nvrtcProgram p = nullptr;
    nvrtcResult r = n.createProgram(&p, src.c_str(), "nb200_custom.cu", 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS) {
        t_log = std::string("nvrtcCreateProgram: ") + n.getErrorString(r);
        return -1;
    }
    // NIT = -1 selects the low-rank engine (nuts_kernel<M, W, 0, true>, lowrank.cuh)
    const std::string nuts_expr = "nb200::nuts_kernel<nb200::CustomModel, " + std::to_string(W) + ", " +
                                  (NIT < 0 ? std::string("0, true") : std::to_string(NIT)) + ">";
    const std::string comp_expr = "nb200::component_kernel<nb200::CustomModel, " + std::to_string(W) + ">";
    n.addNameExpression(p, nuts_expr.c_str());
    n.addNameExpression(p, comp_expr.c_str());
    const std::string dW = "-DNB200_RTC_W=" + std::to_string(W);
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-DNB200_RTC=1", dW.c_str(),
                          "-lineinfo", "-default-device"};
    r = n.compileProgram(p, (int)(sizeof(opts) / sizeof(opts[0])), opts);
    size_t ls = 0;
    n.getProgramLogSize(p, &ls);
    std::string log(ls, '\0');
    if (ls > 1) n.getProgramLog(p, &log[0]);
    if (r != NVRTC_SUCCESS) {
        t_log = std::string("NVRTC: ") + n.getErrorString(r) + "\n" + log;
        n.destroyProgram(&p);
        return -1;
    }
    auto b = std::make_unique<Built>();
    size_t cs = 0;
    n.getCUBINSize(p, &cs);
    b->cubin.resize(cs);
    n.getCUBIN(p, b->cubin.data());
    const char* low = nullptr;
    if (n.getLoweredName(p, nuts_expr.c_str(), &low) == NVRTC_SUCCESS && low) b->nuts_name = low;
    if (n.getLoweredName(p, comp_expr.c_str(), &low) == NVRTC_SUCCESS && low) b->comp_name = low;
    n.destroyProgram(&p);
    if (cs == 0 || b->nuts_name.empty() || b->comp_name.empty()) {
        t_log = "NVRTC produced no cubin / kernel names";
        return -1;
    }
    *out = b.get();
    prog.built[key] = std::move(b);
    return 0;
}

cudaError_t load_locked(Built& b) {
    if (b.lib) return cudaSuccess;
    cudaError_t e = cudaLibraryLoadData(&b.lib, b.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) return e;
    if ((e = cudaLibraryGetKernel(&b.nuts, b.lib, b.nuts_name.c_str())) != cudaSuccess) return e;
    return cudaLibraryGetKernel(&b.comp, b.lib, b.comp_name.c_str());
}

}  // namespace

int custom_register(const char* source) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (size_t i = 0; i < g_programs.size(); ++i)
        if (g_programs[i]->source == source) return (int)i;  // same text: reuse compiled kernels
    g_programs.push_back(std::make_unique<Program>());
    g_programs.back()->source = source;
    return (int)g_programs.size() - 1;
}

int custom_compile(int program, int W, int NIT) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (program < 0 || program >= (int)g_programs.size()) {
        t_log = "unknown custom program";
        return -1;
    }
    Built* b = nullptr;
    return compile_locked(*g_programs[program], W, NIT, &b);
}

const char* custom_last_log() { return t_log.c_str(); }

static cudaError_t get_built(int program, int W, int NIT, Built** out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (program < 0 || program >= (int)g_programs.size()) return cudaErrorInvalidValue;
    if (compile_locked(*g_programs[program], W, NIT, out) != 0) return cudaErrorInvalidSource;
    return load_locked(**out);
}

template <>
cudaError_t launch_nuts<CustomModel>(int W, int NIT, const KParams<CustomModel>& P,
                                     size_t smem_per_chain, size_t block_data, int cpb, int grid,
                                     int block, cudaStream_t stream) {
    Built* b = nullptr;
    cudaError_t e = get_built(P.mdata.program, W, NIT, &b);
    if (e != cudaSuccess) return e;
    const size_t smem = block_data + smem_per_chain * cpb;
    const void* fn = reinterpret_cast<const void*>(b->nuts);
    if ((e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
        return e;
    cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    KParams<CustomModel> Pc = P;
    void* args[] = {&Pc, &smem_per_chain, &block_data};
    return cudaLaunchKernel(fn, dim3(grid), dim3(block), args, smem, stream);
}

template <>
cudaError_t launch_nuts_lr<CustomModel>(int W, const KParams<CustomModel>& P, size_t smem_per_chain,
                                        size_t block_data, int cpb, int grid, int block,
                                        cudaStream_t stream) {
    return launch_nuts<CustomModel>(W, -1, P, smem_per_chain, block_data, cpb, grid, block, stream);
}

template <>
cudaError_t launch_component<CustomModel>(int W, const KParams<CustomModel>& P, int mode,
                                          const double* scal, double* out, size_t smem, unsigned n) {
    Built* b = nullptr;
    cudaError_t e = get_built(P.mdata.program, W, 0, &b);
    if (e != cudaSuccess) return e;
    const void* fn = reinterpret_cast<const void*>(b->comp);
    if ((e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
        return e;
    KParams<CustomModel> Pc = P;
    void* args[] = {&Pc, &mode, &scal, &out};
    return cudaLaunchKernel(fn, dim3(n), dim3(32 * W), args, smem, nullptr);
}

template int supported_nit<CustomModel>(int, int);
template size_t smem_fixed<CustomModel>(int, const CustomModel::Data&, int);
template size_t model_block_data_bytes<CustomModel>(const CustomModel::Data&);
template bool supports_pipeline<CustomModel>(int);
template int sub_warp_lanes<CustomModel>(int, int);
template cudaError_t launch_nuts_sub<CustomModel>(int, int, const KParams<CustomModel>&, size_t, int, int,
                                                  cudaStream_t);
template cudaError_t launch_nuts_piped<CustomModel>(int, const KParams<CustomModel>&, size_t, size_t,
                                                    int, int, cudaStream_t);

}  // namespace nb200
