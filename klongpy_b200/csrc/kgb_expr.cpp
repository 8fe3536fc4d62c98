// Fused expression kernels: postfix program -> CUDA source (macros + jit_skeleton.cuh) -> NVRTC ->
// sm_100a cubin -> cuModule, cached per (device, program).  Replaces the per-node NumPy calls of
// NumpyBackendProvider._ir_to_source (klongpy/backends/numpy_backend.py:116-178) with ONE pass over memory.
#include <dlfcn.h>
#include <nvrtc.h>

#include <cmath>
#include <cstring>
#include <mutex>
#include <sstream>
#include <unordered_map>
#include <vector>

#include "kgb_internal.h"

extern const char* kgb_jit_skeleton_src;  // generated from jit_skeleton.cuh at build time

namespace kgb {

// ------------------------------------------------------------------------------------ driver / NVRTC
struct Driver {
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                             CUstream, void**, void**) = nullptr;
    CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    bool ok = false;
};

static Driver& driver() {
    static Driver d;
    static std::once_flag once;
    std::call_once(once, [] {
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess &&
                   q == cudaDriverEntryPointSuccess && *fn != nullptr;
        };
        bool ok = true;
        ok &= get("cuModuleLoadData", (void**)&d.ModuleLoadData);
        ok &= get("cuModuleGetFunction", (void**)&d.ModuleGetFunction);
        ok &= get("cuLaunchKernel", (void**)&d.LaunchKernel);
        ok &= get("cuOccupancyMaxActiveBlocksPerMultiprocessor",
                  (void**)&d.OccupancyMaxActiveBlocksPerMultiprocessor);
        ok &= get("cuGetErrorString", (void**)&d.GetErrorString);
        ok &= get("cuFuncSetAttribute", (void**)&d.FuncSetAttribute);
        d.ok = ok;
    });
    return d;
}

static const char* cu_err(CUresult r) {
    const char* s = nullptr;
    if (driver().GetErrorString) driver().GetErrorString(r, &s);
    return s ? s : "unknown CUDA driver error";
}

struct Nvrtc {
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*,
                                 const char* const*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    const char* (*GetErrorString)(nvrtcResult) = nullptr;
    bool ok = false;
    std::string why;
};

static Nvrtc& nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("KGB200_NVRTC");
        const char* cands[] = {env,
                               "/usr/local/cuda/lib64/libnvrtc.so.12",
                               "/usr/local/cuda/lib64/libnvrtc.so",
                               "libnvrtc.so.12",
                               "libnvrtc.so"};
        void* h = nullptr;
        for (const char* c : cands) {
            if (!c || !*c) continue;
            h = dlopen(c, RTLD_NOW | RTLD_LOCAL);
            if (h) break;
        }
        if (!h) {
            n.why = "cannot dlopen libnvrtc.so.12 (set KGB200_NVRTC)";
            return;
        }
        bool ok = true;
        auto sym = [&](const char* name, void** fn) {
            *fn = dlsym(h, name);
            ok &= (*fn != nullptr);
        };
        sym("nvrtcCreateProgram", (void**)&n.CreateProgram);
        sym("nvrtcCompileProgram", (void**)&n.CompileProgram);
        sym("nvrtcGetCUBINSize", (void**)&n.GetCUBINSize);
        sym("nvrtcGetCUBIN", (void**)&n.GetCUBIN);
        sym("nvrtcGetProgramLogSize", (void**)&n.GetProgramLogSize);
        sym("nvrtcGetProgramLog", (void**)&n.GetProgramLog);
        sym("nvrtcDestroyProgram", (void**)&n.DestroyProgram);
        sym("nvrtcGetErrorString", (void**)&n.GetErrorString);
        n.ok = ok;
        if (!ok) n.why = "libnvrtc is missing required symbols";
    });
    return n;
}

// ------------------------------------------------------------------------------------ codegen
struct Ty {
    int dt;
    bool weak;  // python scalar / literal: takes part in promotion like NumPy-2 weak scalars
};
struct Val {
    std::string s;
    Ty t;
};

static const char* ctype(int dt) {
    switch (dt) {
        case KGB_F64: return "double";
        case KGB_I64: return "i64";
        case KGB_F32: return "float";
        case KGB_I32: return "int";
        case KGB_B8: return "u8";
    }
    return "?";
}
static bool is_flt(int dt) { return dt == KGB_F64 || dt == KGB_F32; }

// NumPy result_type for the dtypes a KlongPy program can hold (numpy_backend.py:185 keeps only i/f/O)
static int promote_strong(int a, int b) {
    if (a == b) return a;
    if (a == KGB_B8) return b;
    if (b == KGB_B8) return a;
    if (a == KGB_F64 || b == KGB_F64) return KGB_F64;
    if (is_flt(a) != is_flt(b)) return KGB_F64;  // f32 with i32/i64 -> f64
    if (is_flt(a)) return KGB_F64;               // unreachable (f32,f32 handled by a==b)
    return KGB_I64;                              // i32 with i64
}
static Ty promote(Ty a, Ty b) {
    if (a.weak && b.weak) return Ty{(is_flt(a.dt) || is_flt(b.dt)) ? KGB_F64 : KGB_I64, true};
    if (!a.weak && !b.weak) return Ty{promote_strong(a.dt, b.dt), false};
    Ty w = a.weak ? a : b, s = a.weak ? b : a;
    if (is_flt(w.dt)) return Ty{is_flt(s.dt) ? s.dt : KGB_F64, false};
    return Ty{s.dt == KGB_B8 ? KGB_I64 : s.dt, false};
}
static std::string cast(const Val& v, int dt) {
    if (v.t.dt == dt) return v.s;
    if (dt == KGB_B8) return "((u8)((" + v.s + ") != 0))";
    return std::string("((") + ctype(dt) + ")(" + v.s + "))";
}

static std::string lit_f64(double d) {
    char buf[96];
    if (std::isfinite(d)) {
        snprintf(buf, sizeof buf, "(%a)", d);
    } else {
        long long bits;
        memcpy(&bits, &d, 8);
        snprintf(buf, sizeof buf, "__longlong_as_double(%lldLL)", bits);
        if (bits == (long long)0x8000000000000000ULL) snprintf(buf, sizeof buf, "(-0.0)");
    }
    return buf;
}
static std::string lit_i64(long long v) {
    char buf[64];
    if (v == (long long)0x8000000000000000ULL)
        snprintf(buf, sizeof buf, "((i64)(-9223372036854775807LL-1))");
    else
        snprintf(buf, sizeof buf, "((i64)%lldLL)", v);
    return buf;
}

struct ArgInfo {
    int kind, dtype;
};

// Translate one postfix program into a C expression.  `j` is the textual register-slot parameter of the
// KGB_EVAL macro.  Returns false (with error set) on malformed / unsupported programs.
static bool gen_expr(const int32_t* code, int n_code, const std::vector<ArgInfo>& args, bool allow_scanval,
                     int scan_dt, Val* out) {
    std::vector<Val> st;
    auto need = [&](size_t k) {
        if (st.size() < k) {
            set_error("expression program: stack underflow");
            return false;
        }
        return true;
    };
    for (int pc = 0; pc < n_code; pc++) {
        const int op = code[pc];
        switch (op) {
            case KGB_OP_ARG: {
                if (pc + 1 >= n_code) return set_error("truncated ARG"), false;
                const int i = code[++pc];
                if (i < 0 || i >= (int)args.size()) return set_error("ARG index %d out of range", i), false;
                const ArgInfo& a = args[i];
                char buf[48];
                if (a.kind == KGB_ARG_ARRAY) {
                    snprintf(buf, sizeof buf, "a%d[(j)]", i);
                    st.push_back(Val{buf, Ty{a.dtype, false}});
                } else {
                    snprintf(buf, sizeof buf, "s%d", i);
                    st.push_back(Val{buf, Ty{a.dtype, a.kind == KGB_ARG_SCALAR}});
                }
                break;
            }
            case KGB_OP_LIT_I64:
            case KGB_OP_LIT_F64: {
                if (pc + 2 >= n_code) return set_error("truncated literal"), false;
                unsigned long long bits = (unsigned)code[pc + 1] | ((unsigned long long)(unsigned)code[pc + 2] << 32);
                pc += 2;
                if (op == KGB_OP_LIT_I64) {
                    st.push_back(Val{lit_i64((long long)bits), Ty{KGB_I64, true}});
                } else {
                    double d;
                    memcpy(&d, &bits, 8);
                    st.push_back(Val{lit_f64(d), Ty{KGB_F64, true}});
                }
                break;
            }
            case KGB_OP_SCANVAL: {
                if (!allow_scanval) return set_error("SCANVAL outside a scan epilogue"), false;
                st.push_back(Val{"(sv)", Ty{scan_dt, false}});
                break;
            }
            case KGB_OP_ADD:
            case KGB_OP_SUB:
            case KGB_OP_MUL:
            case KGB_OP_DIV:
            case KGB_OP_POW:
            case KGB_OP_MAX:
            case KGB_OP_MIN:
            case KGB_OP_FMOD:
            case KGB_OP_FLOORDIV:
            case KGB_OP_MOD: {
                if (!need(2)) return false;
                Val b = st.back();
                st.pop_back();
                Val a = st.back();
                st.pop_back();
                Ty rt = promote(a.t, b.t);
                if (rt.dt == KGB_B8 && op != KGB_OP_MAX && op != KGB_OP_MIN) rt.dt = KGB_I64;
                if (op == KGB_OP_DIV && !is_flt(rt.dt)) rt.dt = KGB_F64;
                const std::string ca = cast(a, rt.dt), cb = cast(b, rt.dt);
                std::string s;
                switch (op) {
                    case KGB_OP_ADD: s = "(" + ca + " + " + cb + ")"; break;
                    case KGB_OP_SUB: s = "(" + ca + " - " + cb + ")"; break;
                    case KGB_OP_MUL: s = "(" + ca + " * " + cb + ")"; break;
                    case KGB_OP_DIV: s = "(" + ca + " / " + cb + ")"; break;
                    case KGB_OP_POW: s = "kgb_pow(" + ca + ", " + cb + ")"; break;
                    case KGB_OP_MAX: s = "kgb_max(" + ca + ", " + cb + ")"; break;
                    case KGB_OP_MIN: s = "kgb_min(" + ca + ", " + cb + ")"; break;
                    case KGB_OP_FMOD: s = "kgb_fmod(" + ca + ", " + cb + ")"; break;
                    case KGB_OP_FLOORDIV: s = "kgb_floordiv(" + ca + ", " + cb + ")"; break;
                    case KGB_OP_MOD: s = "kgb_mod(" + ca + ", " + cb + ")"; break;
                }
                st.push_back(Val{s, rt});
                break;
            }
            case KGB_OP_EQ:
            case KGB_OP_GT:
            case KGB_OP_LT:
            case KGB_OP_EQ_B:
            case KGB_OP_GT_B:
            case KGB_OP_LT_B:
            case KGB_OP_NE_B:
            case KGB_OP_GE_B:
            case KGB_OP_LE_B: {
                if (!need(2)) return false;
                Val b = st.back();
                st.pop_back();
                Val a = st.back();
                st.pop_back();
                Ty ct = promote(a.t, b.t);
                const std::string ca = cast(a, ct.dt), cb = cast(b, ct.dt);
                const char* c = "==";
                bool as_bool = true;
                switch (op) {
                    case KGB_OP_EQ: c = "=="; as_bool = false; break;
                    case KGB_OP_GT: c = ">"; as_bool = false; break;
                    case KGB_OP_LT: c = "<"; as_bool = false; break;
                    case KGB_OP_EQ_B: c = "=="; break;
                    case KGB_OP_GT_B: c = ">"; break;
                    case KGB_OP_LT_B: c = "<"; break;
                    case KGB_OP_NE_B: c = "!="; break;
                    case KGB_OP_GE_B: c = ">="; break;
                    case KGB_OP_LE_B: c = "<="; break;
                }
                std::string s = std::string("((") + (as_bool ? "u8" : "i64") + ")(" + ca + " " + c + " " + cb + "))";
                st.push_back(Val{s, Ty{as_bool ? KGB_B8 : KGB_I64, false}});
                break;
            }
            case KGB_OP_ISCLOSE: {
                if (!need(2)) return false;
                Val b = st.back();
                st.pop_back();
                Val a = st.back();
                st.pop_back();
                st.push_back(Val{"((u8)kgb_isclose(" + cast(a, KGB_F64) + ", " + cast(b, KGB_F64) + "))",
                                 Ty{KGB_B8, false}});
                break;
            }
            case KGB_OP_WHERE: {
                if (!need(3)) return false;
                Val b = st.back();
                st.pop_back();
                Val a = st.back();
                st.pop_back();
                Val c = st.back();
                st.pop_back();
                Ty rt = promote(a.t, b.t);
                st.push_back(Val{"((" + c.s + ") ? " + cast(a, rt.dt) + " : " + cast(b, rt.dt) + ")", Ty{rt.dt, false}});
                break;
            }
            default: {
                // monads
                if (!need(1)) return false;
                Val a = st.back();
                st.pop_back();
                Ty t = a.t;
                std::string s;
                auto math1 = [&](const char* fn) {
                    const int rt = (t.dt == KGB_F32) ? KGB_F32 : KGB_F64;
                    s = std::string("((") + ctype(rt) + ")" + fn + "(" + cast(a, KGB_F64) + "))";
                    t = Ty{rt, t.weak && is_flt(t.dt)};
                    t.weak = false;
                };
                switch (op) {
                    case KGB_OP_NEG:
                        if (t.dt == KGB_B8) { a.s = cast(a, KGB_I64); t.dt = KGB_I64; }
                        s = "(-" + a.s + ")";
                        break;
                    case KGB_OP_ABS: s = "kgb_abs(" + a.s + ")"; break;
                    case KGB_OP_SQUARE:
                        if (t.dt == KGB_B8) { a.s = cast(a, KGB_I64); t.dt = KGB_I64; }
                        s = "(" + a.s + " * " + a.s + ")";
                        break;
                    case KGB_OP_SIGN:
                        if (t.dt == KGB_B8) { a.s = cast(a, KGB_I64); t.dt = KGB_I64; }
                        s = "kgb_sign(" + a.s + ")";
                        break;
                    case KGB_OP_FLOOR:
                        s = is_flt(t.dt) ? std::string("((") + ctype(t.dt) + ")floor((double)" + a.s + "))" : a.s;
                        break;
                    case KGB_OP_CEIL:
                        s = is_flt(t.dt) ? std::string("((") + ctype(t.dt) + ")ceil((double)" + a.s + "))" : a.s;
                        break;
                    case KGB_OP_TRUNC:
                        s = is_flt(t.dt) ? std::string("((") + ctype(t.dt) + ")trunc((double)" + a.s + "))" : a.s;
                        break;
                    case KGB_OP_FLOOR_I64:
                        s = is_flt(t.dt) ? "((i64)floor((double)" + a.s + "))" : cast(a, KGB_I64);
                        t = Ty{KGB_I64, false};
                        break;
                    case KGB_OP_RECIP:
                        s = "(1.0 / " + cast(a, KGB_F64) + ")";
                        t = Ty{KGB_F64, false};
                        break;
                    case KGB_OP_NOT_B:
                        s = "((u8)((" + a.s + ") == 0))";
                        t = Ty{KGB_B8, false};
                        break;
                    case KGB_OP_ISNAN_B:
                        s = is_flt(t.dt) ? "((u8)((" + a.s + ") != (" + a.s + ")))" : std::string("((u8)0)");
                        t = Ty{KGB_B8, false};
                        break;
                    case KGB_OP_ISINF_B:
                        s = is_flt(t.dt) ? "((u8)isinf((double)" + a.s + "))" : std::string("((u8)0)");
                        t = Ty{KGB_B8, false};
                        break;
                    case KGB_OP_TO_F64: s = cast(a, KGB_F64); t = Ty{KGB_F64, false}; break;
                    case KGB_OP_TO_I64: s = cast(a, KGB_I64); t = Ty{KGB_I64, false}; break;
                    case KGB_OP_TO_F32: s = cast(a, KGB_F32); t = Ty{KGB_F32, false}; break;
                    case KGB_OP_TO_I32: s = cast(a, KGB_I32); t = Ty{KGB_I32, false}; break;
                    case KGB_OP_TO_B8: s = cast(a, KGB_B8); t = Ty{KGB_B8, false}; break;
                    case KGB_OP_SQRT: math1("sqrt"); break;
                    case KGB_OP_EXP: math1("exp"); break;
                    case KGB_OP_LOG: math1("log"); break;
                    case KGB_OP_SIN: math1("sin"); break;
                    case KGB_OP_COS: math1("cos"); break;
                    case KGB_OP_TAN: math1("tan"); break;
                    case KGB_OP_TANH: math1("tanh"); break;
                    case KGB_OP_LOG2: math1("log2"); break;
                    case KGB_OP_LOG10: math1("log10"); break;
                    case KGB_OP_SINH: math1("sinh"); break;
                    case KGB_OP_COSH: math1("cosh"); break;
                    case KGB_OP_ASIN: math1("asin"); break;
                    case KGB_OP_ACOS: math1("acos"); break;
                    case KGB_OP_ATAN: math1("atan"); break;
                    case KGB_OP_EXPM1: math1("expm1"); break;
                    case KGB_OP_LOG1P: math1("log1p"); break;
                    default: return set_error("expression program: unknown opcode %d", op), false;
                }
                st.push_back(Val{s, t});
            }
        }
    }
    if (st.size() != 1) return set_error("expression program: %zu values left on the stack", st.size()), false;
    *out = st.back();
    return true;
}

static int fold_acc_type(int dt, int red_op) {
    // np.add.reduce / np.cumsum promote bool and small ints to int64; max/min keep the dtype
    if (red_op == KGB_RED_ADD || red_op == KGB_RED_MUL) {
        if (dt == KGB_B8 || dt == KGB_I32) return KGB_I64;
    }
    return dt;
}
static std::string fold_macro(const char* name, int red_op) {
    std::string s = std::string("#define ") + name + "(a, b) ";
    switch (red_op) {
        case KGB_RED_ADD: s += "((a) + (b))"; break;
        case KGB_RED_MUL: s += "((a) * (b))"; break;
        case KGB_RED_MAX: s += "kgb_max((a), (b))"; break;
        case KGB_RED_MIN: s += "kgb_min((a), (b))"; break;
    }
    return s + "\n";
}
static std::string fold_ident(const char* name, int red_op, int dt) {
    std::string s = std::string("#define ") + name + " ";
    switch (red_op) {
        case KGB_RED_ADD: s += "0"; break;
        case KGB_RED_MUL: s += "1"; break;
        case KGB_RED_MAX: s += std::string("kgb_lim<") + ctype(dt) + ">::lo()"; break;
        case KGB_RED_MIN: s += std::string("kgb_lim<") + ctype(dt) + ">::hi()"; break;
    }
    return s + "\n";
}

static uint64_t fnv1a(const std::string& s) {
    uint64_t h = 1469598103934665603ULL;
    for (unsigned char c : s) {
        h ^= c;
        h *= 1099511628211ULL;
    }
    return h;
}

}  // namespace kgb

// ------------------------------------------------------------------------------------ handle
struct kgb_expr {
    int device = -1;
    int mode = 0, red_op = 0, red_op2 = 0;
    int out_dtype = 0, acc_dtype = 0, acc2_dtype = 0;
    int vec = 1, align = 16, unroll = 4;
    std::vector<kgb::ArgInfo> args;
    std::string source, name_v, name_s;
    CUmodule module = nullptr;
    CUfunction fn_v = nullptr, fn_s = nullptr;
    int occ_v = 1, occ_s = 1;
    int sblock = 256, sitems = 16;  // scan tile geometry (threads, elements per thread)
    int pipe_arg = -1;              // >= 0: the persistent pipelined scan kernel exists (index of its one array argument)
    std::string name_p;
    CUfunction fn_p = nullptr;
    int pipe_threads = 256, pipe_stages = 7, pipe_load = 2, pipe_park = 4, pipe_lw = 2, pipe_lbw = 1;
    int pipe_ctas_per_sm = 1;
    bool has_red2 = false;          // two-output reduce (second result at out + 8 bytes)
    bool has_epi = false;           // two-output reduce whose last CTA combines both folds into ONE result (kgb_expr_compile_epi)
};

namespace kgb {

struct KgbParamsHost {
    const void* ptr[KGB_MAX_ARGS];
    double sf[KGB_MAX_ARGS];
    long long si[KGB_MAX_ARGS];
    void* out;
    void* ws;
    long long n;
};

static std::mutex g_cache_mu;
static std::unordered_map<std::string, kgb_expr*> g_cache;

// Scan tile geometry.  The chained look-back moves (32-tile window / state round-trip) tiles per unit time, so
// the tile must carry enough bytes for that rate to exceed HBM bandwidth (DESIGN.md "scan"); the tile also has
// to fit shared memory several times per SM.  Overridable for tuning: KGB200_SCAN_BLOCK / KGB200_SCAN_ITEMS.
static void scan_geometry(kgb_expr* h) {
    const int sz = (int)dtype_size(h->acc_dtype);
    h->sblock = 256;
    h->sitems = 128 / sz;  // 128 B per thread
    if (const char* e = getenv("KGB200_SCAN_BLOCK")) {
        const int v = atoi(e);
        if (v >= 32 && v <= 1024 && (v & (v - 1)) == 0) h->sblock = v;
    }
    if (const char* e = getenv("KGB200_SCAN_ITEMS")) {
        const int v = atoi(e) / sz;  // given in bytes per thread
        if (v >= 16 / sz && v * sz <= 512 && (v * sz) % 64 == 0) h->sitems = v;
    }
}

// persistent pipelined scan (jit_skeleton.cuh KGB_PIPE): `stages` x 32 KB + control block; compute threads + look-back warps
static const long long kPipeTile = 4096;
static size_t pipe_smem(const kgb_expr* h) { return (size_t)h->pipe_stages * 32768 + 1024; }
static unsigned pipe_block(const kgb_expr* h) { return (unsigned)(h->pipe_threads + 32 * h->pipe_lw); }
// Geometry: KGB200_PIPE="threads,stages,load,park,lookback_warps,ctas_per_sm" overrides the default for tuning.
static void pipe_geometry(kgb_expr* h) {
    if (const char* e = getenv("KGB200_PIPE_LBW")) {
        const int v = atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8) h->pipe_lbw = v;
    }
    if (const char* e = getenv("KGB200_PIPE")) {
        int t, s, l, d, w, c;
        if (sscanf(e, "%d,%d,%d,%d,%d,%d", &t, &s, &l, &d, &w, &c) == 6 && (t == 256 || t == 512) && s >= l + d + 1 && s <= 7 &&
            l >= 1 && d >= 1 && w >= 1 && w <= 8 && c >= 1 && c <= 2) {
            h->pipe_threads = t, h->pipe_stages = s, h->pipe_load = l, h->pipe_park = d, h->pipe_lw = w, h->pipe_ctas_per_sm = c;
        }
    }
}

// dynamic shared memory of a scan tile: KGB_SBLOCK * (KGB_ITEMS + 16 B pad) accumulators (jit_skeleton.cuh)
static size_t scan_smem_bytes(const kgb_expr* h) {
    const size_t sz = dtype_size(h->acc_dtype);
    return (size_t)h->sblock * ((size_t)h->sitems * sz + 16);
}

static int build_source(kgb_expr* h, const int32_t* code, int n_code, const int32_t* code2, int n_code2,
                        const int32_t* code3 = nullptr, int n_code3 = 0) {
    Val v;
    if (!gen_expr(code, n_code, h->args, false, 0, &v)) return KGB_EINVAL;
    int res_dt = v.t.dt;
    std::ostringstream m;
    m << "#define KGB_MODE " << h->mode << "\n";
    int maxsz = 1;
    for (auto& a : h->args)
        if (a.kind == KGB_ARG_ARRAY) maxsz = std::max(maxsz, (int)dtype_size(a.dtype));

    if (h->mode == KGB_MODE_MAP) {
        h->out_dtype = res_dt;
        h->acc_dtype = res_dt;
    } else {
        h->acc_dtype = fold_acc_type(res_dt, h->red_op);
        h->out_dtype = h->acc_dtype;
    }
    Val v2;
    if (h->mode == KGB_MODE_SCANRED) {
        if (!code2 || n_code2 <= 0) return set_error("SCANRED needs an epilogue program"), KGB_EINVAL;
        if (!gen_expr(code2, n_code2, h->args, true, h->acc_dtype, &v2)) return KGB_EINVAL;
        h->acc2_dtype = fold_acc_type(v2.t.dt, h->red_op2);
        h->out_dtype = h->acc2_dtype;
    }
    maxsz = std::max(maxsz, (int)dtype_size(h->out_dtype));
    maxsz = std::max(maxsz, (int)dtype_size(h->acc_dtype));
    // MAP / REDUCE move 32-byte groups (256-bit LDG/STG, two groups in flight per array per thread); the scan
    // kernels keep 16-byte groups (their shared-memory transposes are built around uint4)
    const bool wide = (h->mode == KGB_MODE_MAP || h->mode == KGB_MODE_REDUCE) && !getenv("KGB200_NO_WIDE");
    h->vec = (wide ? 32 : 16) / maxsz;
    h->align = wide ? 32 : 16;
    h->unroll = wide ? 2 : 4;

    m << "#define KGB_VEC " << h->vec << "\n#define KGB_UNROLL " << h->unroll << "\n";
    if (const char* e = getenv("KGB200_LDHINT")) m << "#define KGB_LDHINT \"" << e << "\"\n#define KGB_STHINT \"" << (getenv("KGB200_STHINT") ? getenv("KGB200_STHINT") : e) << "\"\n";
    m << "#define KGB_OUT_T " << ctype(h->out_dtype) << "\n";
    m << "#define KGB_ACC_T " << ctype(h->acc_dtype) << "\n";
    if (h->mode == KGB_MODE_SCAN || h->mode == KGB_MODE_SCANRED) {
        scan_geometry(h);
        m << "#define KGB_SBLOCK " << h->sblock << "\n#define KGB_ITEMS " << h->sitems << "\n";
    }
    // scalars
    m << "#define KGB_SCALARS";
    for (size_t i = 0; i < h->args.size(); i++) {
        const ArgInfo& a = h->args[i];
        if (a.kind == KGB_ARG_SCALAR) {
            if (a.dtype == KGB_F64)
                m << " const double s" << i << " = p.sf[" << i << "];";
            else
                m << " const i64 s" << i << " = p.si[" << i << "];";
        } else if (a.kind == KGB_ARG_DEVSCALAR) {
            m << " const " << ctype(a.dtype) << " s" << i << " = *(const " << ctype(a.dtype) << "*)p.ptr[" << i << "];";
        }
    }
    m << "\n#define KGB_DECL_REGS(N)";
    for (size_t i = 0; i < h->args.size(); i++)
        if (h->args[i].kind == KGB_ARG_ARRAY) m << " " << ctype(h->args[i].dtype) << " a" << i << "[N];";
    m << "\n#define KGB_LOAD_VEC(V, off, idx)";
    for (size_t i = 0; i < h->args.size(); i++)
        if (h->args[i].kind == KGB_ARG_ARRAY) {
            const char* t = ctype(h->args[i].dtype);
            m << " kgb_ldv<" << t << ", V>(a" << i << " + (off), (const " << t << "*)p.ptr[" << i << "] + (idx));";
        }
    m << "\n#define KGB_EVAL(j) (" << cast(v, h->mode == KGB_MODE_MAP ? h->out_dtype : h->acc_dtype) << ")\n";
    if (h->mode != KGB_MODE_MAP) {
        m << fold_macro("KGB_RED", h->red_op) << fold_ident("KGB_RED_IDENT", h->red_op, h->acc_dtype);
    }
    if (h->mode == KGB_MODE_REDUCE && code2 && n_code2 > 0) {
        // two-output reduce: a second expression over the same arguments folded with red_op2 in the same pass
        Val vb;
        if (!gen_expr(code2, n_code2, h->args, false, 0, &vb)) return KGB_EINVAL;
        h->acc2_dtype = fold_acc_type(vb.t.dt, h->red_op2);
        if (dtype_size(h->acc_dtype) != 8 || dtype_size(h->acc2_dtype) != 8)
            return set_error("two-output reduce needs 8-byte accumulators"), KGB_EINVAL;
        h->has_red2 = true;
        m << "#define KGB_HAS_RED2 1\n#define KGB_ACC2_T " << ctype(h->acc2_dtype) << "\n";
        m << "#define KGB_EVALB(j) (" << cast(vb, h->acc2_dtype) << ")\n";
        m << fold_macro("KGB_RED2", h->red_op2) << fold_ident("KGB_RED2_IDENT", h->red_op2, h->acc2_dtype);
        if (code3 && n_code3 > 0) {
            // epilogue over the two finished folds (`(+/p*s)%+/s`: the divide), evaluated once by the last CTA: the
            // program's argument 0 is the first fold, argument 1 the second; the kernel then has ONE result
            std::vector<ArgInfo> eargs = {ArgInfo{KGB_ARG_DEVSCALAR, h->acc_dtype}, ArgInfo{KGB_ARG_DEVSCALAR, h->acc2_dtype}};
            Val ve;
            if (!gen_expr(code3, n_code3, eargs, false, 0, &ve)) return KGB_EINVAL;
            h->has_epi = true;
            h->out_dtype = ve.t.dt;
            m << "#define KGB_HAS_EPI 1\n#define KGB_EPI_T " << ctype(ve.t.dt) << "\n#define KGB_EPI(s0, s1) (" << ve.s << ")\n";
        }
    } else if (code3 && n_code3 > 0) {
        return set_error("an epilogue program needs a two-output reduce"), KGB_EINVAL;
    }
    if (h->mode == KGB_MODE_SCAN && code2 && n_code2 > 0) {
        // optional seed program over scalar arguments only (multi-GPU scans pass the shard's exclusive offset)
        Val vc;
        if (!gen_expr(code2, n_code2, h->args, false, 0, &vc)) return KGB_EINVAL;
        if (vc.s.find("[(j)]") != std::string::npos) return set_error("scan seed program must not read arrays"), KGB_EINVAL;
        m << "#define KGB_HAS_CARRY 1\n#define KGB_CARRY (" << cast(vc, h->acc_dtype) << ")\n";
    }
    if ((h->mode == KGB_MODE_SCAN || h->mode == KGB_MODE_SCANRED) && dtype_size(h->acc_dtype) == 8 &&
        (h->mode == KGB_MODE_SCANRED || dtype_size(h->out_dtype) == 8) && !getenv("KGB200_NO_PIPE")) {
        // persistent pipelined variant (jit_skeleton.cuh, KGB_PIPE): exactly one array argument, 8-byte elements
        int n_arr = 0, idx = -1;
        for (size_t i = 0; i < h->args.size(); i++)
            if (h->args[i].kind == KGB_ARG_ARRAY) {
                n_arr++;
                idx = (int)i;
            }
        if (n_arr == 1 && dtype_size(h->args[idx].dtype) == 8) {
            h->pipe_arg = idx;
            pipe_geometry(h);
            m << "#define KGB_PT " << h->pipe_threads << "\n#define KGB_PS " << h->pipe_stages << "\n#define KGB_PP " << h->pipe_load
              << "\n#define KGB_PD " << h->pipe_park << "\n#define KGB_PLW " << h->pipe_lw << "\n#define KGB_PMINB " << h->pipe_ctas_per_sm << "\n#define KGB_PLBW " << h->pipe_lbw << "\n";
            m << "#define KGB_PIPE 1\n#define KGB_PIPE_T " << ctype(h->args[idx].dtype) << "\n#define KGB_PIPE_ARG a" << idx
              << "\n#define KGB_PIPE_PTR p.ptr[" << idx << "]\n";
        }
    }
    if (h->mode == KGB_MODE_SCANRED) {
        m << "#define KGB_ACC2_T " << ctype(h->acc2_dtype) << "\n";
        m << "#define KGB_EVAL2(j, sv) (" << cast(v2, h->acc2_dtype) << ")\n";
        m << fold_macro("KGB_RED2", h->red_op2) << fold_ident("KGB_RED2_IDENT", h->red_op2, h->acc2_dtype);
    }
    const std::string macros = m.str();
    char hx[32];
    snprintf(hx, sizeof hx, "%016llx", (unsigned long long)fnv1a(macros));
    static const char* mode_names[] = {"map", "reduce", "scan", "scanred"};
    std::string stem = std::string("kgb_") + mode_names[h->mode] + "_" + hx;
    h->name_v = stem + "_v";
    h->name_s = stem + "_s";
    h->name_p = stem + "_p";
    h->source = "// generated by libkgb200 (kgb_expr.cpp)\n#define KGB_KERNEL_NAME_V " + h->name_v +
                "\n#define KGB_KERNEL_NAME_S " + h->name_s + "\n#define KGB_KERNEL_NAME_P " + h->name_p + "\n" + macros +
                kgb_jit_skeleton_src;
    return KGB_OK;
}

static int jit_compile(kgb_expr* h, bool load) {
    Nvrtc& n = nvrtc();
    if (!n.ok) return set_error("NVRTC unavailable: %s", n.why.c_str()), KGB_ECOMPILE;
    nvrtcProgram prog;
    std::string fname = h->name_v + ".cu";
    nvrtcResult r = n.CreateProgram(&prog, h->source.c_str(), fname.c_str(), 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS) return set_error("nvrtcCreateProgram: %s", n.GetErrorString(r)), KGB_ECOMPILE;
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "-lineinfo",
                          "--extra-device-vectorization"};
    r = n.CompileProgram(prog, 5, opts);
    if (r != NVRTC_SUCCESS) {
        size_t ls = 0;
        n.GetProgramLogSize(prog, &ls);
        std::string log(ls + 1, '\0');
        if (ls) n.GetProgramLog(prog, &log[0]);
        set_error("NVRTC compile failed (%s):\n%s", n.GetErrorString(r), log.c_str());
        n.DestroyProgram(&prog);
        return KGB_ECOMPILE;
    }
    size_t cs = 0;
    n.GetCUBINSize(prog, &cs);
    std::vector<char> cubin(cs);
    n.GetCUBIN(prog, cubin.data());
    n.DestroyProgram(&prog);
    if (const char* dump = getenv("KGB200_DUMP_DIR")) {  // keep source + cubin for cuobjdump / ncu source view
        std::string base = std::string(dump) + "/" + h->name_v;
        if (FILE* f = fopen((base + ".cu").c_str(), "w")) {
            fwrite(h->source.data(), 1, h->source.size(), f);
            fclose(f);
        }
        if (FILE* f = fopen((base + ".cubin").c_str(), "wb")) {
            fwrite(cubin.data(), 1, cubin.size(), f);
            fclose(f);
        }
    }
    if (!load) return KGB_OK;
    Driver& d = driver();
    if (!d.ok) return set_error("CUDA driver entry points unavailable"), KGB_ECUDA;
    CUresult cr = d.ModuleLoadData(&h->module, cubin.data());
    if (cr != CUDA_SUCCESS) return set_error("cuModuleLoadData: %s", cu_err(cr)), KGB_ECUDA;
    cr = d.ModuleGetFunction(&h->fn_v, h->module, h->name_v.c_str());
    if (cr != CUDA_SUCCESS) return set_error("cuModuleGetFunction(%s): %s", h->name_v.c_str(), cu_err(cr)), KGB_ECUDA;
    cr = d.ModuleGetFunction(&h->fn_s, h->module, h->name_s.c_str());
    if (cr != CUDA_SUCCESS) return set_error("cuModuleGetFunction(%s): %s", h->name_s.c_str(), cu_err(cr)), KGB_ECUDA;
    if (h->pipe_arg >= 0) {
        cr = d.ModuleGetFunction(&h->fn_p, h->module, h->name_p.c_str());
        if (cr != CUDA_SUCCESS) return set_error("cuModuleGetFunction(%s): %s", h->name_p.c_str(), cu_err(cr)), KGB_ECUDA;
        cr = d.FuncSetAttribute(h->fn_p, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)pipe_smem(h));
        if (cr != CUDA_SUCCESS) return set_error("cuFuncSetAttribute(%s, %zu B): %s", h->name_p.c_str(), pipe_smem(h), cu_err(cr)), KGB_ECUDA;
    }
    const bool is_scan = h->mode == KGB_MODE_SCAN || h->mode == KGB_MODE_SCANRED;
    const size_t dyn = is_scan ? scan_smem_bytes(h) : 0;
    if (dyn) {
        cr = d.FuncSetAttribute(h->fn_v, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)dyn);
        if (cr == CUDA_SUCCESS) cr = d.FuncSetAttribute(h->fn_s, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)dyn);
        if (cr != CUDA_SUCCESS) return set_error("cuFuncSetAttribute(max dynamic smem %zu): %s", dyn, cu_err(cr)), KGB_ECUDA;
    }
    d.OccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_v, h->fn_v, is_scan ? h->sblock : 256, dyn);
    d.OccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_s, h->fn_s, is_scan ? h->sblock : 256, dyn);
    if (h->occ_v < 1) h->occ_v = 1;
    if (h->occ_s < 1) h->occ_s = 1;
    return KGB_OK;
}

static const int kBlock = 256;

static long long scan_tile(const kgb_expr* h) { return (long long)h->sblock * h->sitems; }

static unsigned fold_grid(const kgb_expr* h, long long n, bool vec) {
    const long long V = vec ? h->vec : 1;
    const long long groups = (n + V - 1) / V;
    long long need = (groups + (long long)kBlock * h->unroll - 1) / ((long long)kBlock * h->unroll);
    long long cap = (long long)sm_count() * (vec ? h->occ_v : h->occ_s);
    // MAP: many waves of CTAs instead of one persistent wave — SMs do not all stream at the same speed, and the
    // hardware CTA scheduler balances them (measured at 2^28 f64: a+b 0.998 -> 1.06 of the copy peak, -a 0.92 -> 1.05;
    // the fold kernels keep one wave: their partial count grows with the grid and they lose 7 % at x64)
    int mult = h->mode == KGB_MODE_MAP ? 32 : 1;
    if (const char* e = getenv("KGB200_GRID_MULT"))
        if (h->mode == KGB_MODE_MAP) mult = std::max(1, atoi(e));  // (the fold workspace is sized for one wave)
    cap *= mult;
    if (need < 1) need = 1;
    return (unsigned)std::min(need, cap);
}

}  // namespace kgb

using namespace kgb;

extern "C" int kgb_expr_compile(const int32_t* code, int32_t n_code, const int32_t* code2, int32_t n_code2,
                                const int32_t* arg_kinds, const int32_t* arg_dtypes, int32_t n_args, int32_t mode,
                                int32_t red_op, int32_t red_op2, kgb_expr** out_handle, int32_t* out_dtype) {
    return kgb_expr_compile_epi(code, n_code, code2, n_code2, nullptr, 0, arg_kinds, arg_dtypes, n_args, mode, red_op, red_op2,
                                out_handle, out_dtype);
}

extern "C" int kgb_expr_compile_epi(const int32_t* code, int32_t n_code, const int32_t* code2, int32_t n_code2,
                                    const int32_t* code3, int32_t n_code3, const int32_t* arg_kinds,
                                    const int32_t* arg_dtypes, int32_t n_args, int32_t mode, int32_t red_op,
                                    int32_t red_op2, kgb_expr** out_handle, int32_t* out_dtype) {
    KGB_REQUIRE(code && n_code > 0 && out_handle, "kgb_expr_compile: null program");
    KGB_REQUIRE(n_args >= 0 && n_args <= KGB_MAX_ARGS, "kgb_expr_compile: %d arguments (max %d)", n_args, KGB_MAX_ARGS);
    KGB_REQUIRE(mode >= KGB_MODE_MAP && mode <= KGB_MODE_SCANRED, "kgb_expr_compile: bad mode %d", mode);
    KGB_REQUIRE(red_op >= 0 && red_op <= 3 && red_op2 >= 0 && red_op2 <= 3, "kgb_expr_compile: bad fold op");
    const int dev = current_device();
    KGB_REQUIRE(dev >= 0, "kgb_init() has not been called on this thread");
    // cache key
    std::string key;
    key.append((const char*)&dev, 4).append((const char*)&mode, 4).append((const char*)&red_op, 4).append((const char*)&red_op2, 4);
    key.append((const char*)&n_args, 4);
    if (n_args) key.append((const char*)arg_kinds, 4 * n_args).append((const char*)arg_dtypes, 4 * n_args);
    key.append((const char*)&n_code, 4).append((const char*)code, 4 * (size_t)n_code);
    if (code2 && n_code2 > 0) key.append((const char*)code2, 4 * (size_t)n_code2);
    if (code3 && n_code3 > 0) key.append("|epi", 4).append((const char*)code3, 4 * (size_t)n_code3);

    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) {
        *out_handle = it->second;
        if (out_dtype) *out_dtype = it->second->out_dtype;
        return KGB_OK;
    }
    kgb_expr* h = new kgb_expr();
    h->device = dev;
    h->mode = mode;
    h->red_op = red_op;
    h->red_op2 = red_op2;
    for (int i = 0; i < n_args; i++) {
        const int k = arg_kinds[i], dt = arg_dtypes[i];
        if (k < 0 || k > 2 || dtype_size(dt) == 0 || (k == KGB_ARG_SCALAR && dt != KGB_F64 && dt != KGB_I64)) {
            delete h;
            set_error("kgb_expr_compile: bad argument %d (kind %d dtype %d)", i, k, dt);
            return KGB_EINVAL;
        }
        h->args.push_back(ArgInfo{k, dt});
    }
    int rc = build_source(h, code, n_code, code2, n_code2, code3, n_code3);
    if (rc == KGB_OK) rc = jit_compile(h, true);
    if (rc != KGB_OK) {
        delete h;
        return rc;
    }
    g_cache.emplace(std::move(key), h);
    *out_handle = h;
    if (out_dtype) *out_dtype = h->out_dtype;
    return KGB_OK;
}

// Code generation + NVRTC compile to an sm_100a cubin WITHOUT loading it: runs on a box with no GPU
// (build check and host-logic tests).  *source_out stays valid until the next call on this thread.
extern "C" int kgb_expr_check(const int32_t* code, int32_t n_code, const int32_t* code2, int32_t n_code2,
                              const int32_t* arg_kinds, const int32_t* arg_dtypes, int32_t n_args, int32_t mode,
                              int32_t red_op, int32_t red_op2, int32_t* out_dtype, const char** source_out) {
    KGB_REQUIRE(code && n_code > 0, "kgb_expr_check: null program");
    KGB_REQUIRE(n_args >= 0 && n_args <= KGB_MAX_ARGS, "kgb_expr_check: %d arguments (max %d)", n_args, KGB_MAX_ARGS);
    KGB_REQUIRE(mode >= KGB_MODE_MAP && mode <= KGB_MODE_SCANRED, "kgb_expr_check: bad mode %d", mode);
    kgb_expr h;
    h.mode = mode;
    h.red_op = red_op;
    h.red_op2 = red_op2;
    for (int i = 0; i < n_args; i++) h.args.push_back(ArgInfo{arg_kinds[i], arg_dtypes[i]});
    int rc = build_source(&h, code, n_code, code2, n_code2);
    if (rc == KGB_OK) rc = jit_compile(&h, false);
    static thread_local std::string last_src;
    last_src = h.source;
    if (source_out) *source_out = last_src.c_str();
    if (out_dtype) *out_dtype = h.out_dtype;
    return rc;
}

extern "C" int kgb_expr_workspace_bytes(const kgb_expr* h, int64_t n, size_t* out_bytes) {
    KGB_REQUIRE(h && out_bytes && n >= 0, "kgb_expr_workspace_bytes: bad arguments");
    size_t b = 0;
    if (h->mode == KGB_MODE_REDUCE) {
        const size_t parts = (size_t)sm_count() * std::max(h->occ_v, h->occ_s);
        b = 64 + align_up(parts * dtype_size(h->acc_dtype), 64);
        if (h->has_red2) b += align_up(parts * dtype_size(h->acc2_dtype), 64);
    } else if (h->mode == KGB_MODE_SCAN || h->mode == KGB_MODE_SCANRED) {
        const long long tl = std::min(scan_tile(h), kPipeTile);
        const size_t tiles = (size_t)((n + tl - 1) / tl);
        b = 64 + tiles * 16;
        if (h->mode == KGB_MODE_SCANRED) b += tiles * dtype_size(h->acc2_dtype);
    }
    *out_bytes = align_up(b, 64);
    return KGB_OK;
}

extern "C" int kgb_expr_launch(kgb_expr* h, const void* const* args, int64_t n, void* out, void* workspace,
                               size_t workspace_bytes, void* stream) {
    return kgb_expr_launch_ex(h, args, n, out, workspace, workspace_bytes, stream, 0);
}

extern "C" int kgb_expr_launch_ex(kgb_expr* h, const void* const* args, int64_t n, void* out, void* workspace,
                                  size_t workspace_bytes, void* stream, int32_t flags) {
    KGB_REQUIRE(h && (out || n == 0), "kgb_expr_launch: null handle / output");
    KGB_REQUIRE(h->device == current_device(), "kgb_expr_launch: handle compiled for device %d, thread bound to %d",
                h->device, current_device());
    KGB_REQUIRE(n >= 0, "kgb_expr_launch: negative length");
    if (n == 0) {
        KGB_REQUIRE(h->mode == KGB_MODE_MAP || h->mode == KGB_MODE_SCAN, "kgb_expr_launch: fold over an empty array");
        return KGB_OK;
    }
    KgbParamsHost P;
    memset(&P, 0, sizeof P);
    const uintptr_t amask = (uintptr_t)h->align - 1;
    bool aligned = ((uintptr_t)out & amask) == 0;
    for (size_t i = 0; i < h->args.size(); i++) {
        const ArgInfo& a = h->args[i];
        KGB_REQUIRE(args && args[i], "kgb_expr_launch: argument %zu is null", i);
        if (a.kind == KGB_ARG_SCALAR) {
            if (a.dtype == KGB_F64)
                memcpy(&P.sf[i], args[i], 8);
            else
                memcpy(&P.si[i], args[i], 8);
        } else {
            P.ptr[i] = args[i];
            if (a.kind == KGB_ARG_ARRAY) aligned &= ((uintptr_t)args[i] & amask) == 0;
        }
    }
    P.out = out;
    P.ws = workspace;
    P.n = n;
    size_t need = 0;
    kgb_expr_workspace_bytes(h, n, &need);
    KGB_REQUIRE(workspace_bytes >= need && (need == 0 || workspace), "kgb_expr_launch: workspace %zu < %zu bytes",
                workspace_bytes, need);
    unsigned grid = 1;
    bool use_pipe = false;
    cudaStream_t st = (cudaStream_t)stream;
    if (h->mode == KGB_MODE_MAP) {
        grid = fold_grid(h, n, aligned);
    } else if (h->mode == KGB_MODE_REDUCE) {
        grid = fold_grid(h, n, aligned);
        // the fold's ticket (ws + 32) must be zero on entry; the kernel's last CTA zeroes it again, so a caller that
        // zero-filled the workspace once and only ever hands it to this library can skip the memset (one driver call)
        if (!(flags & KGB_LAUNCH_WS_CLEAN)) KGB_CUDA_CHECK(cudaMemsetAsync(workspace, 0, 64, st));
    } else {
        const size_t tiles = (size_t)((n + scan_tile(h) - 1) / scan_tile(h));
        KGB_REQUIRE(tiles < 0xffffffffull, "kgb_expr_launch: too many scan tiles");
        grid = (unsigned)tiles;
        const size_t ptiles = (size_t)((n + kPipeTile - 1) / kPipeTile);
        // the persistent variant pays a pipeline fill per CTA: worth it once every SM has a few tiles
        use_pipe = h->fn_p && aligned && ptiles >= (size_t)sm_count() * 4;
        if (use_pipe) grid = (unsigned)std::min<size_t>(ptiles, (size_t)sm_count() * h->pipe_ctas_per_sm);
        KGB_CUDA_CHECK(cudaMemsetAsync(workspace, 0, 64 + std::max(tiles, ptiles) * 16, st));
    }
    if (use_pipe) {
        void* pp[] = {&P};
        CUresult pr = driver().LaunchKernel(h->fn_p, grid, 1, 1, pipe_block(h), 1, 1, (unsigned)pipe_smem(h), (CUstream)stream, pp, nullptr);
        if (pr != CUDA_SUCCESS) return set_error("cuLaunchKernel(%s): %s", h->name_p.c_str(), cu_err(pr)), KGB_ECUDA;
        return KGB_OK;
    }
    void* params[] = {&P};
    const unsigned block = (h->mode == KGB_MODE_SCAN || h->mode == KGB_MODE_SCANRED) ? (unsigned)h->sblock : (unsigned)kBlock;
    const unsigned dyn_smem = (h->mode == KGB_MODE_SCAN || h->mode == KGB_MODE_SCANRED) ? (unsigned)scan_smem_bytes(h) : 0u;
    CUresult cr = driver().LaunchKernel(aligned ? h->fn_v : h->fn_s, grid, 1, 1, block, 1, 1, dyn_smem, (CUstream)stream,
                                        params, nullptr);
    if (cr != CUDA_SUCCESS) return set_error("cuLaunchKernel(%s): %s", h->name_v.c_str(), cu_err(cr)), KGB_ECUDA;
    return KGB_OK;
}

// Result element type of a MAP program for the given argument kinds / dtypes: host-only type inference with the
// same promotion rules the code generator applies (no NVRTC, no GPU) — what a lazy host layer needs to answer
// `.dtype` without launching anything.
extern "C" int kgb_expr_infer(const int32_t* code, int32_t n_code, const int32_t* arg_kinds, const int32_t* arg_dtypes,
                              int32_t n_args, int32_t* out_dtype) {
    KGB_REQUIRE(code && n_code > 0 && out_dtype, "kgb_expr_infer: null program");
    KGB_REQUIRE(n_args >= 0 && n_args <= KGB_MAX_ARGS, "kgb_expr_infer: %d arguments (max %d)", n_args, KGB_MAX_ARGS);
    std::vector<ArgInfo> args;
    for (int i = 0; i < n_args; i++) args.push_back(ArgInfo{arg_kinds[i], arg_dtypes[i]});
    Val v;
    if (!gen_expr(code, n_code, args, false, 0, &v)) return KGB_EINVAL;
    *out_dtype = v.t.dt;
    return KGB_OK;
}

extern "C" int kgb_expr_out_dtype2(const kgb_expr* h) { return (h && h->has_red2 && !h->has_epi) ? h->acc2_dtype : -1; }
extern "C" const char* kgb_expr_source(const kgb_expr* h) { return h ? h->source.c_str() : ""; }
extern "C" const char* kgb_expr_kernel_name(const kgb_expr* h) { return h ? h->name_v.c_str() : ""; }
