// Scene-specialised kernels: the decoded microcode of one program is turned into straight-line
// CUDA C++ (every parameter an immediate, every value slot a register), compiled at run time
// with NVRTC for sm_100a and loaded through the runtime's library API.
//
// This is SURVEY.md §8(f) rank 2 ("scene-specialised kernels via NVRTC"; the reference
// itself can print a straight-line C evaluator: nodes/codegen.py:137-204).  The generated
// code calls the very same op library (cc_ops.cuh / cc_math.cuh) and the same kernel body
// (cc_body.cuh: coordinates, sinks, ordered compaction) as the interpreter, compiled with the
// same -fmad=false contract, so results are bit-identical; what disappears is instruction
// fetch, decode, dispatch, parameter loads and the shared-memory value slots.
//
// NVRTC is loaded with dlopen: without it cc_program_specialize() reports an error and the
// interpreter kernels remain the (only) path.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "cc_internal.h"

namespace {

// ---- NVRTC through dlopen ------------------------------------------------------------------------
typedef struct _nvrtcProgram *nvrtcProgram;
struct Nvrtc {
    void *handle = nullptr;
    int (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *);
    int (*CompileProgram)(nvrtcProgram, int, const char *const *);
    int (*GetProgramLogSize)(nvrtcProgram, size_t *);
    int (*GetProgramLog)(nvrtcProgram, char *);
    int (*GetCUBINSize)(nvrtcProgram, size_t *);
    int (*GetCUBIN)(nvrtcProgram, char *);
    int (*DestroyProgram)(nvrtcProgram *);
    const char *(*GetErrorString)(int);
};

bool load_nvrtc(Nvrtc *n, std::string *err)
{
    static Nvrtc cached;
    if (cached.handle) {
        *n = cached;
        return true;
    }
    const char *names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    void *h = nullptr;
    for (const char *nm : names)
        if ((h = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) break;
    if (!h) {
        *err = "NVRTC (libnvrtc.so.12) not found: scene-specialised kernels unavailable";
        return false;
    }
    Nvrtc r;
    r.handle = h;
#define SYM(field, name)                                                  \
    *(void **)(&r.field) = dlsym(h, name);                                \
    if (!r.field) {                                                       \
        *err = std::string("NVRTC symbol missing: ") + name;              \
        return false;                                                     \
    }
    SYM(CreateProgram, "nvrtcCreateProgram")
    SYM(CompileProgram, "nvrtcCompileProgram")
    SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    SYM(GetProgramLog, "nvrtcGetProgramLog")
    SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    SYM(GetCUBIN, "nvrtcGetCUBIN")
    SYM(DestroyProgram, "nvrtcDestroyProgram")
    SYM(GetErrorString, "nvrtcGetErrorString")
#undef SYM
    cached = r;
    *n = r;
    return true;
}

std::string source_dir()
{
    Dl_info info;
    if (dladdr((void *)&source_dir, &info) && info.dli_fname) {
        std::string p(info.dli_fname);
        size_t k = p.rfind('/');
        return (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/csrc/";
    }
    return "csrc/";
}

bool read_file(const std::string &path, std::string *out)
{
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f) return false;
    std::stringstream ss;
    ss << f.rdbuf();
    *out = ss.str();
    return true;
}

// ---- code generation ---------------------------------------------------------------------------------

struct Gen {
    const std::vector<uint32_t> &code;
    std::ostringstream body, consts;
    int n_tables = 0;

    explicit Gen(const std::vector<uint32_t> &c) : code(c) {}

    std::string F(uint32_t idx) const  // exact fp32 literal of microcode word idx
    {
        char buf[48];
        std::snprintf(buf, sizeof buf, "__uint_as_float(0x%08xu)", code[idx]);
        return buf;
    }
    std::string args(uint32_t pc, int first, int count) const
    {
        std::string s;
        for (int i = 0; i < count; ++i) s += (i ? ", " : "") + F(pc + first + i);
        return s;
    }
    std::string C(uint32_t idx) const  // constant-expression literal (for __constant__ initialisers)
    {
        float f;
        std::memcpy(&f, &code[idx], 4);
        if (f != f) return "__builtin_nanf(\"\")";
        if (f - f != 0.0f) return f > 0 ? "__builtin_huge_valf()" : "-__builtin_huge_valf()";
        char buf[48];
        std::snprintf(buf, sizeof buf, "%af", (double)f);  // C++17 hexadecimal floating literal: exact
        return buf;
    }
    static std::string S(uint32_t slot) { return "S" + std::to_string(slot); }
};

int generate(const cc_decoded &dec, int pts, unsigned sink_mask, std::string *src, std::string *err)
{
    Gen g(dec.microcode);
    const std::vector<uint32_t> &c = dec.microcode;
    std::ostringstream &o = g.body;
    uint32_t pc = 0;
    for (;;) {
        if (pc >= c.size()) {
            *err = "internal: microcode without RETURN";
            return CC_ERR_INVALID_PROGRAM;
        }
        const uint32_t h = c[pc], op = CC_HDR_OP(h), src_slot = CC_HDR_SRC(h), dst = CC_HDR_DST(h);
        const std::string B = Gen::S(src_slot) + "[j]";
        o << "        // pc " << pc << "\n";
        switch (op) {
        case MOP_RETURN: break;
        case MOP_NOP: break;
        case MOP_LOAD: o << "        CC_EACH L[j] = " << B << ";\n"; break;
        case MOP_PRIM_CIRCLE:
        case MOP_PRIM_RECT:
            o << "        { const float m[12] = {" << g.args(pc, 1, 12) << "};\n"
              << "          const float mf[12] = {" << g.args(pc, 17, 10) << ", 0.f, 0.f};\n"
              << "          cc_prim_n<" << (op == MOP_PRIM_RECT ? "true" : "false") << ", PTS>(m, mf, " << g.args(pc, 13, 4)
              << ", gx, gy, gz, L); }\n";
            break;
        case MOP_RECTANGLE: o << "        cc_rectangle_n<PTS>(" << g.args(pc, 1, 2) << ", L);\n"; break;
        case MOP_CIRCLE: o << "        cc_circle_n<PTS>(" << g.F(pc + 1) << ", L);\n"; break;
        case MOP_SPHERE: o << "        cc_sphere_n<PTS>(" << g.F(pc + 1) << ", L);\n"; break;
        case MOP_REGPOLY: o << "        CC_EACH L[j] = cc_regular_polygon2d(" << g.args(pc, 1, 5) << ", L[j]);\n"; break;
        case MOP_POLYGON: {
            const uint32_t n = (uint32_t)(*reinterpret_cast<const float *>(&c[pc + 1]));
            const uint32_t off = c[pc + 2];
            const int k = g.n_tables++;
            g.consts << "__constant__ float cc_poly_" << k << "[" << 6 * n << "] = {";
            for (uint32_t i = 0; i < 6 * n; ++i) g.consts << (i ? ", " : "") << g.C(off + i);
            g.consts << "};\n";
            o << "        CC_EACH L[j] = cc_polygon2d_table(cc_poly_" << k << ", " << n << "u, L[j]);\n";
            break;
        }
        case MOP_HALF_SPACE: o << "        CC_EACH L[j] = make_float4(0.0f, -1.0f, 0.0f, -L[j].y);\n"; break;
        case MOP_REV_TO:
            o << "        CC_EACH L[j] = make_float4(cc_len2(L[j].x, L[j].z), L[j].y, 0.0f, 0.0f);\n";
            break;
        case MOP_TWIST_TO: o << "        CC_EACH L[j] = cc_twist_revolution_to(" << g.args(pc, 1, 2) << ", L[j]);\n"; break;
        case MOP_T_INIT:
            o << "        { const float m[12] = {" << g.args(pc, 1, 12) << "};\n"
              << "          CC_EACH L[j] = cc_transform(m, gx[j], gy[j], gz[j]); }\n";
            break;
        case MOP_T_TO:
            o << "        { const float m[12] = {" << g.args(pc, 1, 12) << "};\n"
              << "          CC_EACH L[j] = cc_transform(m, L[j].x, L[j].y, L[j].z); }\n";
            break;
        case MOP_T_FROM:
            o << "        { const float m[12] = {" << g.args(pc, 1, 10) << ", 0.f, 0.f};\n"
              << "          CC_EACH L[j] = cc_transform_from(m, L[j]); }\n";
            break;
        case MOP_MIRROR: o << "        CC_EACH L[j].x = -L[j].x;\n"; break;
        case MOP_SYM_TO: o << "        CC_EACH L[j].x = fabsf(L[j].x);\n"; break;
        case MOP_OFFSET: o << "        CC_EACH L[j].w = L[j].w - " << g.F(pc + 1) << ";\n"; break;
        case MOP_SHELL:
            o << "        CC_EACH { float4 s = (L[j].w >= 0.0f) ? L[j] : cc_neg4(L[j]); s.w = s.w - " << g.F(pc + 1)
              << "; L[j] = s; }\n";
            break;
        case MOP_REPETITION:
            o << "        CC_EACH L[j] = make_float4(cc_remainder(L[j].x, " << g.F(pc + 1) << "), cc_remainder(L[j].y, "
              << g.F(pc + 2) << "), cc_remainder(L[j].z, " << g.F(pc + 3) << "), 0.0f);\n";
            break;
        case MOP_CREP_TO: o << "        CC_EACH L[j] = cc_circular_repetition_to(" << g.args(pc, 1, 2) << ", L[j]);\n"; break;
        case MOP_CREP_FROM:
            o << "        CC_EACH L[j] = cc_circular_repetition_from(" << g.args(pc, 1, 2) << ", L[j], " << B << ");\n";
            break;
        case MOP_GEAR: o << "        CC_EACH L[j] = cc_involute_gear(" << g.args(pc, 1, 5) << ", L[j]);\n"; break;
        case MOP_EXTRUSION:
            o << "        { float cz[PTS]; CC_EACH cz[j] = " << B << ".z; cc_extrusion_n<PTS>(" << g.F(pc + 1)
              << ", L, cz); }\n";
            break;
        case MOP_REV_FROM: o << "        CC_EACH L[j] = cc_revolution_from(L[j], " << B << ");\n"; break;
        case MOP_TWIST_FROM:
            o << "        CC_EACH L[j] = cc_twist_revolution_from(" << g.args(pc, 1, 5) << ", L[j], " << B << ");\n";
            break;
        case MOP_SYM_FROM: o << "        CC_EACH L[j].x = (" << B << ".x < 0.0f) ? -L[j].x : L[j].x;\n"; break;
        case MOP_UNION: o << "        CC_EACH { const float4 b = " << B << "; L[j] = (L[j].w < b.w) ? L[j] : b; }\n"; break;
        case MOP_UNION_R: o << "        CC_EACH L[j] = cc_rounded_union(" << g.F(pc + 1) << ", L[j], " << B << ");\n"; break;
        case MOP_ISECT:
            o << "        CC_EACH { const float4 b = " << B << "; L[j] = (-L[j].w < -b.w) ? L[j] : b; }\n";
            break;
        case MOP_ISECT_R:
            o << "        CC_EACH L[j] = cc_neg4(cc_rounded_union(" << g.F(pc + 1) << ", cc_neg4(L[j]), cc_neg4(" << B
              << ")));\n";
            break;
        case MOP_SUB:
            o << "        CC_EACH { const float4 b = " << B << "; L[j] = (-L[j].w < b.w) ? L[j] : cc_neg4(b); }\n";
            break;
        case MOP_SUB_R:
            o << "        CC_EACH L[j] = cc_neg4(cc_rounded_union(" << g.F(pc + 1) << ", cc_neg4(L[j]), " << B << "));\n";
            break;
        default:
            *err = "internal: unknown micro-op " + std::to_string(op);
            return CC_ERR_INVALID_PROGRAM;
        }
        if (op == MOP_RETURN) break;
        if (dst != CC_SLOT_NONE) o << "        CC_EACH " << Gen::S(dst) << "[j] = L[j];\n";
        pc += CC_HDR_LEN(h);
    }

    std::ostringstream s;
    s << "// generated by libcodecad_b200 (cc_jit.cpp) from " << dec.info.n_micro_ops << " micro-ops\n"
      << "#include \"cc_ops.cuh\"\n#include \"cc_body.cuh\"\n"
      << "#define PTS " << pts << "\n"
      << "#define CC_EACH _Pragma(\"unroll\") for (int j = 0; j < PTS; ++j)\n"
      << g.consts.str() << "struct SceneEval {\n"
      << "    __device__ __forceinline__ void operator()(const float (&gx)[PTS], const float (&gy)[PTS],\n"
      << "                                               const float (&gz)[PTS], float4 (&L)[PTS]) const\n    {\n";
    for (uint32_t k = 0; k < dec.info.n_slots; ++k) s << "        float4 S" << k << "[PTS];\n";
    s << "        CC_EACH L[j] = make_float4(0.f, 0.f, 0.f, 0.f);\n" << o.str() << "    }\n};\n";
    const char *names[4] = {"float4", "pymcubes", "classify", "mass"};
    const char *sinks[4] = {"CC_SINK_FLOAT4", "CC_SINK_PYMCUBES", "CC_SINK_CLASSIFY", "CC_SINK_MASS"};
    // optional occupancy hint (tuning experiments): CODECAD_B200_JIT_MINB = min CTAs per SM
    std::string bounds = "CC_THREADS";
    if (const char *mb = getenv("CODECAD_B200_JIT_MINB")) bounds += std::string(", ") + std::to_string(atoi(mb));
    for (int k = 0; k < 4; ++k)
        if (sink_mask & (1u << k))
        s << "extern \"C\" __global__ void __launch_bounds__(" << bounds << ") cc_jit_" << names[k]
          << "(const cc_eval_args a)\n{\n    SceneEval e;\n    cc_kernel_body<PTS, " << sinks[k] << ">(a, e);\n}\n";
    *src = s.str();
    return CC_OK;
}

}  // namespace

void cc_jit_release(cc_program *prog);

int cc_jit_nvrtc(const std::string &src, std::vector<char> *cubin, std::string *err)
{
    Nvrtc n;
    if (!load_nvrtc(&n, err)) return CC_ERR_CUDA;
    const char *hdr_names[] = {"cc_device_types.h", "cc_math.cuh", "cc_ops.cuh", "cc_body.cuh"};
    std::string hdr_src[4];
    const std::string dir = source_dir();
    for (int i = 0; i < 4; ++i)
        if (!read_file(dir + hdr_names[i], &hdr_src[i])) {
            *err = "cannot read " + dir + hdr_names[i] + " (needed to specialise kernels)";
            return CC_ERR_INVALID_ARGUMENT;
        }
    const char *hdr_ptrs[4] = {hdr_src[0].c_str(), hdr_src[1].c_str(), hdr_src[2].c_str(), hdr_src[3].c_str()};
    nvrtcProgram p = nullptr;
    int e = n.CreateProgram(&p, src.c_str(), "cc_scene.cu", 4, hdr_ptrs, hdr_names);
    if (e) {
        *err = std::string("nvrtcCreateProgram: ") + n.GetErrorString(e);
        return CC_ERR_CUDA;
    }
    const char *opts[] = {"--gpu-architecture=sm_100a", "-fmad=false", "--std=c++17", "-lineinfo"};
    e = n.CompileProgram(p, 4, opts);
    if (e) {
        size_t ls = 0;
        n.GetProgramLogSize(p, &ls);
        std::string log(ls, '\0');
        if (ls) n.GetProgramLog(p, &log[0]);
        if (log.size() > 2000) log.resize(2000);
        *err = std::string("nvrtcCompileProgram: ") + n.GetErrorString(e) + "\n" + log;
        n.DestroyProgram(&p);
        return CC_ERR_CUDA;
    }
    size_t sz = 0;
    n.GetCUBINSize(p, &sz);
    cubin->resize(sz);
    n.GetCUBIN(p, cubin->data());
    n.DestroyProgram(&p);
    return CC_OK;
}

int cc_jit_compile(cc_program *prog, int pts, unsigned sink_mask, double *seconds, std::string *err)
{
    if (pts != 1 && pts != 2 && pts != 4) pts = 2;
    if ((sink_mask & 15u) == 0) sink_mask = 15u;
    auto t0 = std::chrono::steady_clock::now();
    std::string src;
    int rc = generate(prog->dec, pts, sink_mask, &src, err);
    if (rc) return rc;
    std::vector<char> cubin;
    rc = cc_jit_nvrtc(src, &cubin, err);
    if (rc) return rc;
    const size_t sz = cubin.size();
    cc_jit_release(prog);

    cudaLibrary_t lib = nullptr;
    cudaError_t ce = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (ce != cudaSuccess) {
        *err = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(ce);
        return CC_ERR_CUDA;
    }
    const char *names[4] = {"cc_jit_float4", "cc_jit_pymcubes", "cc_jit_classify", "cc_jit_mass"};
    for (int k = 0; k < 4; ++k) {
        if (!(sink_mask & (1u << k))) continue;
        cudaKernel_t kern = nullptr;
        ce = cudaLibraryGetKernel(&kern, lib, names[k]);
        if (ce != cudaSuccess) {
            *err = std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(ce);
            cudaLibraryUnload(lib);
            return CC_ERR_CUDA;
        }
        prog->jit_kernel[k] = (void *)kern;
    }
    prog->jit_library = (void *)lib;
    prog->jit_pts = pts;
    prog->jit_cubin_bytes = sz;
    prog->use_jit = true;
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return CC_OK;
}

void cc_jit_release(cc_program *prog)
{
    if (prog->jit_library) cudaLibraryUnload((cudaLibrary_t)prog->jit_library);
    prog->jit_library = nullptr;
    for (int k = 0; k < 4; ++k) prog->jit_kernel[k] = nullptr;
}

int cc_jit_launch(const cc_program *prog, int sink, const cc_eval_args &a, void *stream)
{
    const uint32_t grid = a.n_blocks * a.tiles_per_block;
    if (grid == 0) return 0;
    void *args[] = {(void *)&a};
    return (int)cudaLaunchKernel((const void *)prog->jit_kernel[sink], dim3(grid), dim3(CC_THREADS), args, 0,
                                 (cudaStream_t)stream);
}

int cc_jit_source(const cc_decoded &dec, int pts, unsigned sink_mask, std::string *src, std::string *err)
{
    if (pts != 1 && pts != 2 && pts != 4) pts = 2;
    if ((sink_mask & 15u) == 0) sink_mask = 15u;
    return generate(dec, pts, sink_mask, src, err);
}
