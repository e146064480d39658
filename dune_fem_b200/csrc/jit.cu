// jit.cu -- generic integrands compiled at run time (include/b200fem.h: b200fem_operator_create_jit).
//
// The reference generates a C++ `Integrands` class from the UFL form and JIT-compiles a GalerkinOperator specialised for
// it (python/dune/models/integrands/model.py:10-106, dune/fem/schemes/integrands.hh:152-375).  The device analogue: the
// user's interior / skeleton / boundary functions (CUDA C++ source text) are wrapped in an Integrands class and NVRTC
// instantiates the generic quadrature kernel (dg_quadrature.cuh -- the very header the built-in integrands are compiled
// from, read from the source tree next to the library) for it: sm_100a cubin -> cuModuleLoadData -> cuLaunchKernel on the
// context's stream.  One compiled kernel per (order, interior rule, surface rule), cached per operator.
// NVRTC and the driver API are bound at run time (dlopen / cudaGetDriverEntryPoint): no link-time dependency.
#include <dlfcn.h>

#include <map>
#include <memory>
#include <mutex>
#include <tuple>

#include "internal.hpp"
#include "jit_integrands.cuh"
#include "lagrange_quadrature.cuh"
#include "launch_dgq.hpp"

namespace b200fem {

namespace {

struct NvrtcApi {
  void* handle = nullptr;
  int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*DestroyProgram)(void**) = nullptr;
  int (*CompileProgram)(void*, int, const char* const*) = nullptr;
  int (*GetProgramLogSize)(void*, size_t*) = nullptr;
  int (*GetProgramLog)(void*, char*) = nullptr;
  int (*GetCUBINSize)(void*, size_t*) = nullptr;
  int (*GetCUBIN)(void*, char*) = nullptr;
  int (*AddNameExpression)(void*, const char*) = nullptr;
  int (*GetLoweredName)(void*, const char*, const char**) = nullptr;
  bool load() {
    if (handle) return true;
    for (const char* n : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"}) { handle = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (handle) break; }
    if (!handle) return false;
    auto sym = [&](const char* n) { return dlsym(handle, n); };
    CreateProgram = (decltype(CreateProgram))sym("nvrtcCreateProgram"); DestroyProgram = (decltype(DestroyProgram))sym("nvrtcDestroyProgram");
    CompileProgram = (decltype(CompileProgram))sym("nvrtcCompileProgram"); GetProgramLogSize = (decltype(GetProgramLogSize))sym("nvrtcGetProgramLogSize");
    GetProgramLog = (decltype(GetProgramLog))sym("nvrtcGetProgramLog"); GetCUBINSize = (decltype(GetCUBINSize))sym("nvrtcGetCUBINSize");
    GetCUBIN = (decltype(GetCUBIN))sym("nvrtcGetCUBIN"); AddNameExpression = (decltype(AddNameExpression))sym("nvrtcAddNameExpression");
    GetLoweredName = (decltype(GetLoweredName))sym("nvrtcGetLoweredName");
    if (!(CreateProgram && DestroyProgram && CompileProgram && GetProgramLogSize && GetProgramLog && GetCUBINSize && GetCUBIN && AddNameExpression && GetLoweredName)) { handle = nullptr; return false; }
    return true;
  }
};
NvrtcApi g_nvrtc;

struct DriverApi {
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
  bool ok = false;
  bool load() {
    if (ok) return true;
    auto get = [](const char* n) { void* fn = nullptr; cudaDriverEntryPointQueryResult q; if (cudaGetDriverEntryPoint(n, &fn, cudaEnableDefault, &q) != cudaSuccess) fn = nullptr; return fn; };
    ModuleLoadData = (decltype(ModuleLoadData))get("cuModuleLoadData"); ModuleUnload = (decltype(ModuleUnload))get("cuModuleUnload");
    ModuleGetFunction = (decltype(ModuleGetFunction))get("cuModuleGetFunction"); FuncSetAttribute = (decltype(FuncSetAttribute))get("cuFuncSetAttribute");
    LaunchKernel = (decltype(LaunchKernel))get("cuLaunchKernel");
    ok = ModuleLoadData && ModuleUnload && ModuleGetFunction && FuncSetAttribute && LaunchKernel;
    return ok;
  }
};
DriverApi g_drv;

// directory of the device headers: <dir of libb200fem.so>/../csrc
std::string csrc_dir() {
  Dl_info info;
  if (!dladdr((const void*)&csrc_dir, &info) || !info.dli_fname) return "";
  std::string p(info.dli_fname); const size_t s = p.rfind('/');
  return (s == std::string::npos ? std::string(".") : p.substr(0, s)) + "/../csrc";
}

// which kernel the integrands are compiled into
enum JitVariant { kJitDg = 0, kJitLagrange3d = 1, kJitLagrange2d = 2, kJitUnstructured2d = 3, kJitUnstructured3d = 4 };

std::string program_text(const std::string& user, bool skel, bool bnd, int N, int MI, int MS, int R, int variant, std::string* name_expr) {
  const std::string rs = std::to_string(R);
  // scalar spaces: PointValue / PointRange; range-R spaces: VectorValue = PointValueV<R>, VectorRange = PointRangeV<R>
  const std::string V = R == 1 ? "PointValue" : "PointValueV<" + rs + ">", G = R == 1 ? "PointRange" : "PointRangeV<" + rs + ">";
  std::string t;
  t += "#include \"lagrange_quadrature.cuh\"\n#include \"lagrange_unstructured.cuh\"\n#include \"jit_integrands.cuh\"\n";
  t += "namespace b200fem {\nnamespace user {\nconstexpr int dimRange = " + rs + ";\nusing VectorValue = PointValueV<dimRange>;\nusing VectorRange = PointRangeV<dimRange>;\n#line 1 \"integrands\"\n" + user + "\n}  // namespace user\n";
  t += "struct JitIntegrands : JitIntegrandsBase {\n"
       "  __device__ static void zero(" + G + "& r) { double* p = reinterpret_cast<double*>(&r); for (int i = 0; i < (int)(sizeof(" + G + ") / sizeof(double)); ++i) p[i] = 0; }\n"
       "  __device__ " + G + " interior(const double* x, const " + V + "& v) const { " + G + " r; zero(r); user::interior(x, v, r, c, dim); return r; }\n"
       "  __device__ void skeleton(const double* x, int axis, double sign, double ihe, const " + V + "& in, const " + V + "& out, " + G + "& rin, " + G + "& rout) const {\n"
       "    zero(rin); zero(rout);\n";
  if (skel) t += "    user::skeleton(x, axis, sign, ihe, in, out, rin, rout, c, dim);\n";
  t += "  }\n  __device__ " + G + " boundary(int axis, int side, double ihbnd, const double* x, const " + V + "& v) const { " + G + " r; zero(r);\n";
  if (bnd) t += "    user::boundary(x, axis, side, ihbnd, v, r, c, dim);\n";
  t += "    return r; }\n};\n}  // namespace b200fem\n";
  const std::string n = std::to_string(N);
  if (variant == kJitDg) *name_expr = "b200fem::dg_quadrature_kernel<" + n + ", " + std::to_string(MI) + ", " + std::to_string(MS) + ", b200fem::JitIntegrands, true, " + rs + ">";
  else if (variant == kJitLagrange3d) *name_expr = "b200fem::lagrange3d_quadrature_kernel<" + n + ", b200fem::JitIntegrands, " + rs + ">";
  else if (variant == kJitLagrange2d) *name_expr = "b200fem::lagrange2d_quadrature_kernel<" + n + ", b200fem::JitIntegrands, " + rs + ">";
  else *name_expr = std::string("b200fem::lagrange_unstructured_kernel<") + (variant == kJitUnstructured2d ? "2, " + std::to_string(N * N) : "3, " + std::to_string(N * N * N)) + ", b200fem::JitIntegrands>";
  return t;
}

// NVRTC: program text -> cubin + lowered kernel name.  No device needed.
int compile(const std::string& user, bool skel, bool bnd, int N, int MI, int MS, int R, int variant, std::vector<char>* cubin, std::string* lowered, std::string* log) {
  if (!g_nvrtc.load()) { if (log) *log = "libnvrtc.so.12 could not be loaded"; return B200FEM_ERR_NOT_IMPLEMENTED; }
  std::string expr; const std::string text = program_text(user, skel, bnd, N, MI, MS, R, variant, &expr);
  void* prog = nullptr;
  if (g_nvrtc.CreateProgram(&prog, text.c_str(), "b200fem_jit.cu", 0, nullptr, nullptr) != 0) { if (log) *log = "nvrtcCreateProgram failed"; return B200FEM_ERR_CUDA; }
  g_nvrtc.AddNameExpression(prog, expr.c_str());
  const std::string inc = "--include-path=" + csrc_dir();
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", inc.c_str()};
  const int rc = g_nvrtc.CompileProgram(prog, 4, opts);
  size_t ls = 0; g_nvrtc.GetProgramLogSize(prog, &ls);
  if (log) { log->assign(ls, '\0'); if (ls) g_nvrtc.GetProgramLog(prog, &(*log)[0]); while (!log->empty() && log->back() == '\0') log->pop_back(); }
  if (rc != 0) { g_nvrtc.DestroyProgram(&prog); return B200FEM_ERR_INVALID; }
  const char* low = nullptr;
  if (g_nvrtc.GetLoweredName(prog, expr.c_str(), &low) != 0 || !low) { g_nvrtc.DestroyProgram(&prog); if (log) *log += "\nno lowered name for " + expr; return B200FEM_ERR_INVALID; }
  if (lowered) *lowered = low;
  size_t cs = 0; g_nvrtc.GetCUBINSize(prog, &cs);
  if (cubin) { cubin->resize(cs); if (cs) g_nvrtc.GetCUBIN(prog, cubin->data()); }
  g_nvrtc.DestroyProgram(&prog);
  return cs ? B200FEM_OK : B200FEM_ERR_INVALID;
}

struct JitKernel { CUmodule mod = nullptr; CUfunction fn = nullptr; };

}  // namespace

struct JitState {
  std::string source; bool skel = false, bnd = false;
  double c[kJitMaxConstants] = {}; int nc = 0;
  std::map<std::tuple<int, int, int, int>, JitKernel> kernels;      // (N, MI, MS, variant)
  double* d_l0 = nullptr; unsigned long long l0_version = ~0ull;     // L[0] for apply_linear, tied to the operator's state version
  double* d_zero = nullptr;
};

void jit_free(b200fem_operator* op) {
  if (!op->jit) return;
  for (auto& kv : op->jit->kernels) if (kv.second.mod && g_drv.ok) g_drv.ModuleUnload(kv.second.mod);
  if (op->jit->d_l0) cudaFree(op->jit->d_l0);
  if (op->jit->d_zero) cudaFree(op->jit->d_zero);
  delete op->jit; op->jit = nullptr;
}

// the compiled kernel of one (order, rules, variant) configuration: NVRTC on first use, cached per operator
// Compiled code is cached per process under everything the program text depends on: a second operator over the same form (another
// mesh, another space of the same order, another rank-local box) loads the cubin instead of running NVRTC again -- the device analogue
// of the reference's on-disk cache of generated operator modules (python/dune/generator).
struct JitCubin { std::vector<char> cubin; std::string lowered; };
static std::map<std::string, JitCubin>& jit_cache() { static std::map<std::string, JitCubin> c; return c; }
static std::mutex g_jit_cache_mutex;

static int jit_kernel(b200fem_operator* op, int N, int MI, int MS, int variant, size_t smem, JitKernel** out) {
  JitState* J = op->jit;
  JitKernel& K = J->kernels[std::make_tuple(N, MI, MS, variant)];
  if (!K.fn) {
    REQUIRE(!op->capturing, B200FEM_ERR_INVALID, "run-time compilation inside a graph capture (apply once before solving)");
    REQUIRE(g_drv.load(), B200FEM_ERR_CUDA, "driver entry points (cuModuleLoadData, cuLaunchKernel) unavailable");
    std::vector<char> cubin; std::string lowered, log;
    const std::string key = std::to_string(N) + "/" + std::to_string(MI) + "/" + std::to_string(MS) + "/" + std::to_string(op->sp->dim_range) + "/" + std::to_string(variant) +
                            (J->skel ? "/s" : "/-") + (J->bnd ? "b/" : "-/") + J->source;
    bool cached = false;
    { std::lock_guard<std::mutex> lock(g_jit_cache_mutex); auto it = jit_cache().find(key); if (it != jit_cache().end()) { cubin = it->second.cubin; lowered = it->second.lowered; cached = true; } }
    int rc = cached ? B200FEM_OK : compile(J->source, J->skel, J->bnd, N, MI, MS, op->sp->dim_range, variant, &cubin, &lowered, &log);
    if (rc) return fail(rc, "integrands do not compile:\n" + log);
    if (!cached) { std::lock_guard<std::mutex> lock(g_jit_cache_mutex); jit_cache()[key] = JitCubin{cubin, lowered}; }
    if (g_drv.ModuleLoadData(&K.mod, cubin.data()) != CUDA_SUCCESS) return fail(B200FEM_ERR_CUDA, "cuModuleLoadData failed for the compiled integrands");
    if (g_drv.ModuleGetFunction(&K.fn, K.mod, lowered.c_str()) != CUDA_SUCCESS) return fail(B200FEM_ERR_CUDA, "compiled kernel not found in its module");
    if (smem && g_drv.FuncSetAttribute(K.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem) != CUDA_SUCCESS) return fail(B200FEM_ERR_CUDA, "cuFuncSetAttribute(max dynamic shared memory) failed");
  }
  *out = &K; return B200FEM_OK;
}
static JitIntegrandsBase jit_params(b200fem_operator* op, const BoxDev& b) {
  JitIntegrandsBase I; std::memset(&I, 0, sizeof(I)); I.m = op->model; I.dim = b.dim; I.with_data = 1; std::memcpy(I.c, op->jit->c, sizeof(I.c));
  return I;
}
// slots of a CTA per element: the components of a range-R space sit in RS = R rounded up to a power of two neighbouring slots
static int slots_per_element(int R) { int p = 1; while (p < R) p *= 2; return p; }

template <int N, int MI, int MS> static int launch_jit_t(b200fem_operator* op, const double* u, double* w, const double* sub) {
  using Cfg = DgQuadCfg<N, MI, MS>; b200fem_ctx* ctx = op->sp->mesh->ctx;
  const BoxDev& b = op->active_box ? *op->active_box : op->sp->box;
  const int epb = Cfg::EB / slots_per_element(op->sp->dim_range);
  REQUIRE(epb >= 1, B200FEM_ERR_NOT_IMPLEMENTED, "compiled integrands: dimRange too large for this order");
  JitKernel* K = nullptr; int rc = jit_kernel(op, N, MI, MS, kJitDg, Cfg::smem_bytes(), &K); if (rc) return rc;
  long long n_owned = (long long)(b.own_hi[0] - b.own_lo[0]) * (b.own_hi[1] - b.own_lo[1]) * (b.own_hi[2] - b.own_lo[2]);
  const unsigned grid = (unsigned)((n_owned + epb - 1) / epb);
  auto tab = make_quad_tab<N, MI, MS>(true, N - 1);
  BoxDev box = b;
  JitIntegrandsBase I = jit_params(op, b);
  const int* perm = op->d_perm; int nbs = op->sp->nb; double scale = mass_scale(op);
  void* args[] = {&tab, &box, &I, &perm, &nbs, &u, &w, &sub, &n_owned, &scale};
  if (g_drv.LaunchKernel(K->fn, grid, 1, 1, Cfg::kThreads, 1, 1, (unsigned)Cfg::smem_bytes(), (CUstream)ctx->stream, args, nullptr) != CUDA_SUCCESS)
    return fail(B200FEM_ERR_CUDA, "cuLaunchKernel failed for the compiled integrands");
  op->timing.launches_per_apply = 1;
  return B200FEM_OK;
}

// continuous Lagrange spaces: w.clear(), then one launch per colour (2^dim colours, plain read-modify-write, lagrange_quadrature.cuh)
template <int N> static int launch_jit_lagrange_t(b200fem_operator* op, const double* u, double* w) {
  using Cfg = DgQuadCfg<N, N, N>; b200fem_space* s = op->sp; b200fem_ctx* ctx = s->mesh->ctx; cudaStream_t st = ctx->stream;
  const BoxDev& b = op->active_box ? *op->active_box : s->box;
  CUDA_OK(cudaMemsetAsync(w, 0, sizeof(double) * (size_t)s->size, st));
  BoxDev box = b; JitIntegrandsBase I = jit_params(op, b); LagrangeLayoutDev L = s->lay;
  int launches = 1;
  if (b.dim == 3) {
    const int epb = Cfg::EB / slots_per_element(s->dim_range);
    REQUIRE(epb >= 1, B200FEM_ERR_NOT_IMPLEMENTED, "compiled integrands: dimRange too large for this order");
    JitKernel* K = nullptr; int rc = jit_kernel(op, N, N, N, kJitLagrange3d, Cfg::smem_bytes(), &K); if (rc) return rc;
    QuadTabDev<N, N, N> T; const Tab1D& t = s->tab;
    for (int i = 0; i < N * N; ++i) { T.Bi[i] = T.Bs[i] = t.B[i]; T.Gi[i] = T.Gs[i] = t.G[i]; }
    for (int i = 0; i < N; ++i) { T.xi[i] = T.xs[i] = t.x[i]; T.wi[i] = T.ws[i] = t.w[i]; T.phi[0][i] = t.phi0[i]; T.phi[1][i] = t.phi1[i]; T.dphi[0][i] = t.dphi0[i]; T.dphi[1][i] = t.dphi1[i]; }
    for (int c = 0; c < 8; ++c) {
      int c0 = c & 1, c1 = (c >> 1) & 1, c2 = c >> 2;
      int m0 = (b.n[0] - c0 + 1) / 2, m1 = (b.n[1] - c1 + 1) / 2, m2 = (b.n[2] - c2 + 1) / 2;
      long long nc = (long long)m0 * m1 * m2; if (nc <= 0) continue;
      void* args[] = {&T, &box, &I, &L, &u, &w, &c0, &c1, &c2, &m0, &m1, &nc};
      if (g_drv.LaunchKernel(K->fn, (unsigned)((nc + epb - 1) / epb), 1, 1, Cfg::kThreads, 1, 1, (unsigned)Cfg::smem_bytes(), (CUstream)st, args, nullptr) != CUDA_SUCCESS)
        return fail(B200FEM_ERR_CUDA, "cuLaunchKernel failed for the compiled integrands");
      ++launches;
    }
  } else {
    JitKernel* K = nullptr; int rc = jit_kernel(op, N, N, N, kJitLagrange2d, 0, &K); if (rc) return rc;
    DgTabDev<N> T; const Tab1D& t = s->tab;
    for (int i = 0; i < N * N; ++i) { T.B[i] = t.B[i]; T.G[i] = t.G[i]; }
    for (int i = 0; i < N; ++i) { T.x[i] = t.x[i]; T.w[i] = t.w[i]; T.phi[0][i] = t.phi0[i]; T.phi[1][i] = t.phi1[i]; T.dphi[0][i] = t.dphi0[i]; T.dphi[1][i] = t.dphi1[i]; }
    for (int c = 0; c < 4; ++c) {
      int c0 = c & 1, c1 = c >> 1; int m0 = (b.n[0] - c0 + 1) / 2, m1 = (b.n[1] - c1 + 1) / 2;
      long long nc = (long long)m0 * m1; if (nc <= 0) continue;
      void* args[] = {&T, &box, &I, &L, &u, &w, &c0, &c1, &m0, &nc};
      if (g_drv.LaunchKernel(K->fn, (unsigned)((nc + 127) / 128), 1, 1, 128, 1, 1, 0, (CUstream)st, args, nullptr) != CUDA_SUCCESS)
        return fail(B200FEM_ERR_CUDA, "cuLaunchKernel failed for the compiled integrands");
      ++launches;
    }
  }
  op->timing.launches_per_apply = launches;
  return B200FEM_OK;
}

// unstructured cube meshes: w.clear(), then one launch per colour of lagrange_unstructured_kernel (index arrays, per-element geometry)
static int launch_jit_unstructured(b200fem_operator* op, const double* u, double* w) {
  b200fem_space* s = op->sp; cudaStream_t st = s->mesh->ctx->stream; const int dim = s->mesh->dim, N = s->n1;
  UnstructuredLaunch L; int rc = unstructured_launch_info(s, &L); if (rc) return rc;
  JitKernel* K = nullptr; rc = jit_kernel(op, N, N, N, dim == 2 ? kJitUnstructured2d : kJitUnstructured3d, L.smem, &K); if (rc) return rc;
  CUDA_OK(cudaMemsetAsync(w, 0, sizeof(double) * (size_t)s->size, st));
  BoxDev b; std::memset(&b, 0, sizeof(b)); b.dim = dim; JitIntegrandsBase I = jit_params(op, b);
  int launches = 1;
  for (size_t c = 0; c + 1 < L.colour_begin->size(); ++c) {
    int first = (*L.colour_begin)[c], count = (*L.colour_begin)[c + 1] - first; if (count <= 0) continue;
    void* args[] = {&L.tab, &I, &L.order, &L.dofs, &L.elem_x, &u, &w, &first, &count};
    if (g_drv.LaunchKernel(K->fn, (unsigned)std::min((count + L.eb - 1) / L.eb, L.max_grid), 1, 1, (unsigned)L.threads, 1, 1, (unsigned)L.smem, (CUstream)st, args, nullptr) != CUDA_SUCCESS)
      return fail(B200FEM_ERR_CUDA, "cuLaunchKernel failed for the compiled integrands");
    ++launches;
  }
  op->timing.launches_per_apply = launches;
  return B200FEM_OK;
}

static int launch_jit(b200fem_operator* op, const double* u, double* w, const double* sub) {
  const int k = op->sp->order, N = op->sp->n1;
  if (op->sp->unst) {
    REQUIRE(default_quadrature(op), B200FEM_ERR_NOT_IMPLEMENTED, "Lagrange spaces: only quadrature orders that select the (order+1)-point Gauss rule");
    int rc = launch_jit_unstructured(op, u, w); if (rc) return rc;
    if (sub) { rc = b200fem_axpy_dev(op, -1.0, sub, w); if (rc) return rc; op->timing.launches_per_apply += 1; }   // L[u] - L[0]
    return B200FEM_OK;
  }
  if (op->sp->kind == B200FEM_LAGRANGE) {
    REQUIRE(default_quadrature(op), B200FEM_ERR_NOT_IMPLEMENTED, "Lagrange spaces: only quadrature orders that select the (order+1)-point Gauss rule");
    REQUIRE(!op->jit->skel, B200FEM_ERR_NOT_IMPLEMENTED, "skeleton integrands on continuous spaces");
    int rc = N == 2 ? launch_jit_lagrange_t<2>(op, u, w) : N == 3 ? launch_jit_lagrange_t<3>(op, u, w) : launch_jit_lagrange_t<4>(op, u, w); if (rc) return rc;
    if (sub) { rc = b200fem_axpy_dev(op, -1.0, sub, w); if (rc) return rc; op->timing.launches_per_apply += 1; }   // L[u] - L[0]
    return B200FEM_OK;
  }
  int mi, ms;
  try { mi = gauss_points_for_order(op->q_interior ? (int)op->q_interior : 2 * k); ms = gauss_points_for_order(op->q_surface ? (int)op->q_surface : 2 * k + 1); }
  catch (const std::exception& ex) { return fail(B200FEM_ERR_NOT_IMPLEMENTED, ex.what()); }
#define B200FEM_JIT_CASE(n)                                                                        \
  if (N == n) {                                                                                    \
    if (mi == n && ms == n) return launch_jit_t<n, n, n>(op, u, w, sub);                           \
    if (mi == n + 1 && ms == n + 1) return launch_jit_t<n, n + 1, n + 1>(op, u, w, sub);           \
    if (mi == n + 1 && ms == n) return launch_jit_t<n, n + 1, n>(op, u, w, sub);                   \
    if (mi == n && ms == n + 1) return launch_jit_t<n, n, n + 1>(op, u, w, sub);                   \
    if (mi == n + 2 && ms == n + 2) return launch_jit_t<n, n + 2, n + 2>(op, u, w, sub);           \
  }
  B200FEM_JIT_CASE(2) B200FEM_JIT_CASE(3) B200FEM_JIT_CASE(4) B200FEM_JIT_CASE(5) B200FEM_JIT_CASE(6)
#undef B200FEM_JIT_CASE
  return fail(B200FEM_ERR_NOT_IMPLEMENTED, "compiled integrands: no kernel configuration for this order / pair of Gauss rules");
}

// w = L[u] (linear == false) or L[u] - L[0] (linear == true)
int apply_jit(b200fem_operator* op, const double* u, double* w, bool linear) {
  JitState* J = op->jit; b200fem_space* s = op->sp; cudaStream_t st = s->mesh->ctx->stream;
  const double* sub = nullptr;
  if (linear) {
    if (!J->d_l0 || J->l0_version != op->state_version) {
      REQUIRE(!op->capturing, B200FEM_ERR_INVALID, "L[0] must exist before graph capture (apply_linear once before solving)");
      const size_t bytes = sizeof(double) * (size_t)s->size;
      if (!J->d_l0) CUDA_OK(cudaMalloc(&J->d_l0, bytes));
      if (!J->d_zero) { CUDA_OK(cudaMalloc(&J->d_zero, bytes)); CUDA_OK(cudaMemsetAsync(J->d_zero, 0, bytes, st)); }
      const BoxDev* ab = op->active_box; op->active_box = nullptr;
      const int rc = launch_jit(op, J->d_zero, J->d_l0, nullptr); op->active_box = ab; if (rc) return rc;
      J->l0_version = op->state_version;
    }
    sub = J->d_l0;
  }
  int rc = launch_jit(op, u, w, sub); if (rc) return rc;
  op->timing.kernel = B200FEM_KERNEL_QUADRATURE;
  return B200FEM_OK;
}

}  // namespace b200fem

using namespace b200fem;

extern "C" int b200fem_jit_compile_check(const char* source, int order, char* log, int log_len) {
  REQUIRE(source && order >= 1 && order <= 5, B200FEM_ERR_INVALID, "jit_compile_check: bad argument");
  std::string lg; const int n = order + 1;
  const int rc = compile(source, true, true, n, n, n, 1, kJitDg, nullptr, nullptr, &lg);
  if (log && log_len > 0) { std::strncpy(log, lg.c_str(), (size_t)log_len - 1); log[log_len - 1] = '\0'; }
  return rc ? fail(rc, lg) : B200FEM_OK;
}

/* the same check for any space the integrands can run on: space kind (b200fem_space_kind), mesh dimension (selects the Lagrange
 * kernel; DG spaces of either dimension share one) and dimRange */
extern "C" int b200fem_jit_compile_check_space(const char* source, int kind, int dim, int order, int dim_range, int has_skeleton, int has_boundary, char* log, int log_len) {
  REQUIRE(source && order >= 1 && order <= 5 && dim_range >= 1 && dim_range <= 4 && (dim == 2 || dim == 3), B200FEM_ERR_INVALID, "jit_compile_check_space: bad argument");
  REQUIRE(kind != B200FEM_LAGRANGE || order <= 3, B200FEM_ERR_NOT_IMPLEMENTED, "Lagrange spaces: orders 1 to 3");
  std::string lg; const int n = order + 1;
  const int variant = kind != B200FEM_LAGRANGE ? kJitDg : dim == 3 ? kJitLagrange3d : kJitLagrange2d;
  const int rc = compile(source, has_skeleton != 0, has_boundary != 0, n, n, n, dim_range, variant, nullptr, nullptr, &lg);
  if (log && log_len > 0) { std::strncpy(log, lg.c_str(), (size_t)log_len - 1); log[log_len - 1] = '\0'; }
  return rc ? fail(rc, lg) : B200FEM_OK;
}

/* ... and for a Lagrange space on an unstructured cube mesh (interior integrands only) */
extern "C" int b200fem_jit_compile_check_unstructured(const char* source, int dim, int order, char* log, int log_len) {
  REQUIRE(source && (order == 1 || order == 2) && (dim == 2 || dim == 3), B200FEM_ERR_INVALID, "jit_compile_check_unstructured: bad argument");
  std::string lg; const int n = order + 1;
  const int rc = compile(source, false, false, n, n, n, 1, dim == 2 ? kJitUnstructured2d : kJitUnstructured3d, nullptr, nullptr, &lg);
  if (log && log_len > 0) { std::strncpy(log, lg.c_str(), (size_t)log_len - 1); log[log_len - 1] = '\0'; }
  return rc ? fail(rc, lg) : B200FEM_OK;
}

extern "C" int b200fem_operator_set_constants(b200fem_operator* op, const double* constants, int nconstants) {
  REQUIRE(op && op->jit, B200FEM_ERR_INVALID, "set_constants: not an operator with compiled integrands");
  REQUIRE(nconstants >= 0 && nconstants <= kJitMaxConstants && (constants || nconstants == 0), B200FEM_ERR_INVALID, "set_constants: at most 32 constants");
  std::memset(op->jit->c, 0, sizeof(op->jit->c));
  for (int i = 0; i < nconstants; ++i) op->jit->c[i] = constants[i];
  op->jit->nc = nconstants;
  invalidate_cached_state(op);
  return B200FEM_OK;
}

extern "C" int b200fem_operator_create_jit(b200fem_space* space, const char* source, const double* constants, int nconstants,
                                           int has_skeleton, int has_boundary, b200fem_operator** out) {
  REQUIRE(space && source && out, B200FEM_ERR_INVALID, "operator_create_jit: null argument");
  REQUIRE(!(space->unst && (has_skeleton || has_boundary)), B200FEM_ERR_NOT_IMPLEMENTED, "unstructured meshes: interior integrands only");
  REQUIRE(!(space->kind == B200FEM_LAGRANGE && has_skeleton), B200FEM_ERR_NOT_IMPLEMENTED, "skeleton integrands on continuous spaces");
  REQUIRE(nconstants >= 0 && nconstants <= kJitMaxConstants, B200FEM_ERR_INVALID, "operator_create_jit: at most 32 constants");
  b200fem_model m; std::memset(&m, 0, sizeof(m)); m.has_skeleton = has_skeleton != 0; m.has_boundary = has_boundary != 0;
  m.gamma = 1.0;     // "not known to be linear": keeps every Kronecker shortcut away from this operator
  b200fem_operator* op = nullptr;
  int rc = operator_create_impl(space, &m, &op); if (rc) return rc;
  op->jit = new JitState; op->jit->source = source; op->jit->skel = m.has_skeleton; op->jit->bnd = m.has_boundary;
  rc = b200fem_operator_set_constants(op, constants, nconstants);
  if (rc) { b200fem_operator_destroy(op); return rc; }
  *out = op; return B200FEM_OK;
}
