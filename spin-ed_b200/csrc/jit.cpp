// jit.cpp -- run-time specialisation of the matvec kernel on the symmetry group.
//
// The group program of a basis (permprog.h) is emitted as straight-line CUDA C++ -- every masked
// rotate becomes funnel shifts and LOP3s with immediate operands, no loop, no loads -- and
// compiled for sm_100a with NVRTC together with matvec_kernel.cuh (the very same device code the
// static kernels use).  The resulting cubin is loaded with cudaLibraryLoadData and launched with
// cudaLaunchKernel.  NVRTC is bound with dlopen; when it is missing, or SPED_JIT=0, the statically
// compiled kernel with the interpreted program runs instead (same algorithm, ~2x the instructions).
#include <dlfcn.h>
#include <nvrtc.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <future>
#include <map>
#include <memory>
#include <mutex>
#include <vector>
#include <sstream>
#include <thread>
#include <tuple>

#include "internal.h"

namespace sped {

extern char const k_src_device_types[];
extern char const k_src_matvec_kernel[];

namespace {

struct NvrtcApi {
  void* handle = nullptr;
  bool tried = false;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
  nvrtcResult (*Version)(int*, int*) = nullptr;
};

NvrtcApi& nvrtc() {
  static NvrtcApi a;
  if (a.tried) return a;
  a.tried = true;
  char const* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                         "/usr/local/cuda/lib64/libnvrtc.so"};
  for (char const* n : names) {
    a.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (a.handle) break;
  }
  if (!a.handle) return a;
  bool ok = true;
  auto load = [&](char const* sym) {
    void* p = dlsym(a.handle, sym);
    if (!p) ok = false;
    return p;
  };
  a.CreateProgram = reinterpret_cast<decltype(a.CreateProgram)>(load("nvrtcCreateProgram"));
  a.CompileProgram = reinterpret_cast<decltype(a.CompileProgram)>(load("nvrtcCompileProgram"));
  a.GetCUBINSize = reinterpret_cast<decltype(a.GetCUBINSize)>(load("nvrtcGetCUBINSize"));
  a.GetCUBIN = reinterpret_cast<decltype(a.GetCUBIN)>(load("nvrtcGetCUBIN"));
  a.GetProgramLogSize = reinterpret_cast<decltype(a.GetProgramLogSize)>(load("nvrtcGetProgramLogSize"));
  a.GetProgramLog = reinterpret_cast<decltype(a.GetProgramLog)>(load("nvrtcGetProgramLog"));
  a.DestroyProgram = reinterpret_cast<decltype(a.DestroyProgram)>(load("nvrtcDestroyProgram"));
  a.Version = reinterpret_cast<decltype(a.Version)>(load("nvrtcVersion"));
  if (!ok) a.handle = nullptr;
  return a;
}

std::string hex32(u32 v) {
  char buf[16];
  std::snprintf(buf, sizeof buf, "0x%08xu", v);
  return buf;
}

// ---- code generation -------------------------------------------------------------------------
// 64-bit images are kept as (lo, hi) 32-bit halves so that rotates are explicit funnel shifts.
struct Gen {
  HostProgram const& P;
  std::ostringstream o;
  bool w32;
  unsigned shift, top;
  u64 all;
  // "top-aligned" form (64-bit words carried with a tag, i.e. P.shift != 0): the word is re-based
  // so that its top spin sits at bit 63.  The spin-inversion test is then the sign of the high half
  // (flipped = hi >> 31, mask = hi >>s 31: two shifts instead of shift + and + negate), and bits
  // outside the spins may hold garbage between steps, which turns every two-displacement rotate
  // into one bit-select (LOP3 with an immediate) per half instead of two; the garbage is cleared
  // by the AND that forms the key.  Measured in SASS on 6x6: 19.5 -> 15 integer instructions per
  // group element; these kernels are bound by the integer pipe.  (A running minimum on doubles was
  // tried and dropped: sm_100 has no DMNMX, fmin becomes DSETP + FSEL + SEL + NaN fix-up.)
  bool dkey;

  explicit Gen(HostProgram const& p) : P(p) {
    w32 = P.word_bits == 32;
    shift = w32 ? 0 : P.shift;
    dkey = false;
    if (!w32 && P.shift != 0 && P.n_spins <= 48 && 2 * P.steps.size() < (1ull << (64 - P.n_spins)) && dkey_enabled()) {
      dkey = true;
      shift = 64 - P.n_spins;
    }
    top = P.n_spins - 1 + shift;
    all = (P.n_spins == 64 ? ~0ull : ((1ull << P.n_spins) - 1)) << shift;
  }
  static bool dkey_enabled() {
    char const* e = std::getenv("SPED_JIT_TOPALIGN");
    return !(e && e[0] == '0');
  }
  // masks of the compiled program are positioned for P.shift; re-base them to `shift`
  u64 rebase(u64 mask) const { return w32 ? mask : ((mask >> P.shift) << shift); }

  // expression for half `h` (0 = lo, 1 = hi) of rotl64((lo,hi), r)
  static std::string rot64(unsigned r, int h) {
    r &= 63u;
    if (r == 0) return h ? "hi" : "lo";
    if (r == 32) return h ? "lo" : "hi";
    char const* A = r < 32 ? "lo" : "hi";  // the pair being shifted is (A = low, B = high)
    char const* B = r < 32 ? "hi" : "lo";
    unsigned s = r & 31u;
    std::ostringstream e;
    if (h) e << "__funnelshift_l(" << A << ", " << B << ", " << s << ")";
    else e << "__funnelshift_l(" << B << ", " << A << ", " << s << ")";
    return e.str();
  }
  static std::string rot32(unsigned r) {
    r &= 31u;
    if (r == 0) return "lo";
    std::ostringstream e;
    e << "__funnelshift_l(lo, lo, " << r << ")";
    return e.str();
  }

  // y <- OR_j (rotl(y, r_j) & m_j)
  void emit_rotmask(std::vector<std::pair<unsigned, u64>> const& terms_in) {
    std::vector<std::pair<unsigned, u64>> terms = terms_in;
    for (auto& t : terms) t.second = rebase(t.second);
    o << "  {\n";
    if (dkey && terms.size() == 1 && terms[0].second == all) {  // pure rotation: nothing to mask
      o << "    u32 nlo = " << rot64(terms[0].first, 0) << ";\n    u32 nhi = " << rot64(terms[0].first, 1)
        << ";\n    hi = nhi;\n    lo = nlo;\n  }\n";
      return;
    }
    if (dkey && terms.size() == 2 && (terms[0].second | terms[1].second) == all && !(terms[0].second & terms[1].second)) {
      // destination bits of the first displacement from rotation 0, everything else from rotation 1;
      // written b ^ ((a ^ b) & M) so that each half is one LOP3 with the immediate M
      u32 mlo = (u32)terms[0].second, mhi = (u32)(terms[0].second >> 32);
      auto select = [&](int h, u32 m) {
        std::string a = rot64(terms[0].first, h), b = rot64(terms[1].first, h);
        if (m == 0u) return b;
        if (m == ~0u) return a;
        return "(" + b + " ^ ((" + a + " ^ " + b + ") & " + hex32(m) + "))";
      };
      o << "    u32 nlo = " << select(0, mlo) << ";\n";
      o << "    u32 nhi = " << select(1, mhi) << ";\n";
      o << "    hi = nhi;\n    lo = nlo;\n  }\n";
      return;
    }
    std::string lo_expr, hi_expr;
    for (size_t j = 0; j < terms.size(); ++j) {
      u32 mlo = (u32)terms[j].second, mhi = (u32)(terms[j].second >> 32);
      if (w32) {
        if (mlo) lo_expr += (lo_expr.empty() ? "" : " | ") + ("(" + rot32(terms[j].first) + " & " + hex32(mlo) + ")");
      } else {
        if (mlo) lo_expr += (lo_expr.empty() ? "" : " | ") + ("(" + rot64(terms[j].first, 0) + " & " + hex32(mlo) + ")");
        if (mhi) hi_expr += (hi_expr.empty() ? "" : " | ") + ("(" + rot64(terms[j].first, 1) + " & " + hex32(mhi) + ")");
      }
    }
    o << "    u32 nlo = " << (lo_expr.empty() ? "0u" : lo_expr) << ";\n";
    if (!w32) o << "    u32 nhi = " << (hi_expr.empty() ? "0u" : hi_expr) << ";\n    hi = nhi;\n";
    o << "    lo = nlo;\n  }\n";
  }

  void emit_benes(std::vector<std::pair<unsigned, u64>> const& swaps_in) {
    std::vector<std::pair<unsigned, u64>> swaps = swaps_in;
    for (auto& t : swaps) t.second = rebase(t.second);
    if (w32) {
      for (auto const& s : swaps)
        o << "  { u32 t = ((lo >> " << s.first << ") ^ lo) & " << hex32((u32)s.second) << "; lo ^= t ^ (t << " << s.first
          << "); }\n";
    } else {
      for (auto const& s : swaps)
        o << "  { u64 y = ((u64)hi << 32) | lo; u64 t = ((y >> " << s.first << ") ^ y) & 0x" << std::hex << s.second
          << std::dec << "ull; y ^= t ^ (t << " << s.first << "); lo = (u32)y; hi = (u32)(y >> 32); }\n";
    }
  }

  // fold the spin inversion into (zlo, zhi, f), form the key and update the running minimum
  void emit_visit(unsigned k) {
    bool inv = P.inversion != 0;
    o << "  {\n";
    if (dkey) {
      if (inv) {
        o << "    u32 f = hi >> 31;\n";                  // the top spin (bit 63) is up: take the inverted image
        o << "    u32 m = (u32)((int)hi >> 31);\n";
        o << "    u32 zhi = (hi ^ m) & " << hex32((u32)(all >> 32)) << ";\n";
        o << "    u32 zlo = ((lo ^ m) & " << hex32((u32)all) << ") | f | " << 2 * k << "u;\n";
      } else {
        o << "    u32 zhi = hi & " << hex32((u32)(all >> 32)) << ";\n";
        o << "    u32 zlo = (lo & " << hex32((u32)all) << ") | " << 2 * k << "u;\n";
      }
      o << "    u64 key = ((u64)zhi << 32) | (u64)zlo;\n";
      o << "    best = key < best ? key : best;\n  }\n";
      return;
    }
    if (inv) {
      if (w32 || top < 32) o << "    u32 f = (lo >> " << top << ") & 1u;\n";
      else o << "    u32 f = (hi >> " << (top - 32) << ") & 1u;\n";
      o << "    u32 m = 0u - f;\n";
      o << "    u32 zlo = lo ^ (" << hex32((u32)all) << " & m);\n";
      if (!w32) o << "    u32 zhi = hi ^ (" << hex32((u32)(all >> 32)) << " & m);\n";
    } else {
      o << "    u32 zlo = lo;\n";
      if (!w32) o << "    u32 zhi = hi;\n";
    }
    std::string tag = std::to_string(2 * k) + "u" + (inv ? " + f" : "");
    if (w32) {
      o << "    u64 key = ((u64)zlo << 32) | (u64)(" << tag << ");\n";
      o << "    best = key < best ? key : best;\n";
    } else if (shift) {
      o << "    u64 key = ((u64)zhi << 32) | (u64)(zlo | (" << tag << "));\n";
      o << "    best = key < best ? key : best;\n";
    } else {
      o << "    u64 z = ((u64)zhi << 32) | zlo;\n";
      o << "    if (z < best) { best = z; bidx = " << tag << "; }\n";
    }
    o << "  }\n";
  }

  std::string run() {
    o << "namespace sped {\n";
    o << "__device__ const int sped_jit_phase[" << P.phase.size() << "] = {";
    for (size_t i = 0; i < P.phase.size(); ++i) o << (i ? "," : "") << P.phase[i];
    o << "};\n";
    o << "__device__ __forceinline__ void sped_jit_canonicalize(u64 x, u64& rep, int& phase) {\n";
    if (w32) o << "  u32 lo = (u32)x;\n";
    else o << "  u64 xs = x << " << shift << ";\n  u32 lo = (u32)xs, hi = (u32)(xs >> 32);\n";
    o << "  u64 best = ~(u64)0;\n";
    if (!w32 && !shift) o << "  u32 bidx = 0;\n";
    emit_visit(0);
    for (size_t k = 1; k < P.steps.size(); ++k) {
      auto const& f = P.fast[k];
      if (!(f.ctl & kFastGeneral)) {
        unsigned r1 = f.ctl & 63u, r2 = (f.ctl >> 8) & 63u;
        std::vector<std::pair<unsigned, u64>> t;
        u64 const all_prog = w32 ? all : ((all >> shift) << P.shift);  // program coordinates; emit_* re-base
        if (r1 == r2) t.push_back({r1, all_prog});
        else {
          t.push_back({r1, f.mask & all_prog});
          t.push_back({r2, ~f.mask & all_prog});
        }
        emit_rotmask(t);
      } else {
        auto const& st = P.steps[k];
        std::vector<std::pair<unsigned, u64>> t;
        for (unsigned j = 0; j < st.n_ops; ++j) t.push_back({P.ops[st.first_op + j].amount, P.ops[st.first_op + j].mask});
        if (st.kind == 0) emit_rotmask(t);
        else emit_benes(t);
      }
      emit_visit((unsigned)k);
    }
    if (w32) {
      o << "  rep = (u64)(u32)(best >> 32);\n  u32 idx = (u32)best;\n";
    } else if (shift) {
      o << "  rep = best >> " << shift << ";\n  u32 idx = (u32)best & " << ((1u << shift) - 1u) << "u;\n";
    } else {
      o << "  rep = best;\n  u32 idx = bidx;\n";
    }
    o << "  int ph = sped_jit_phase[idx >> 1];\n";
    if (P.inversion < 0) o << "  if (idx & 1u) { ph += " << P.denom / 2 << "; if (ph >= " << P.denom << ") ph -= " << P.denom << "; }\n";
    o << "  phase = ph;\n}\n}  // namespace sped\n";
    return o.str();
  }
};

// A compilation in flight: NVRTC runs on its own thread (it needs neither the device nor the basis
// object -- only the generated source), started when the basis is built, so that it overlaps the
// enumeration of the representatives instead of preceding the first application of an operator.
struct Pending {
  std::shared_future<std::vector<char>> cubin;
  std::chrono::steady_clock::time_point started;
  double expected_seconds = 0;  // rough: 0.25 s + 4 ms per program step on the hosts measured
};

// one loaded module = one kernel: the matrix-free matvec of a (dtype, columns) pair, or the cache fill
struct Entry {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kernel = nullptr;
  bool failed = false;
  std::shared_ptr<Pending> pending;
};
constexpr int kKindMatvec = 0, kKindFill = 1;

struct Cache {
  std::map<std::tuple<int, int, int>, Entry> entries;  // (kind, dtype, columns)
  ~Cache() {
    for (auto& e : entries) {
      // a compilation still in flight is waited for: once the last handle is gone the process may
      // exit, and NVRTC must not be running while its own statics are torn down
      if (e.second.pending) e.second.pending->cubin.wait();
      if (e.second.lib) cudaLibraryUnload(e.second.lib);
    }
  }
};

std::mutex g_jit_mutex;

char const* dtype_name(int dtype) {
  switch (dtype) {
    case SPED_F32: return "float";
    case SPED_F64: return "double";
    case SPED_C64: return "float2";
    default: return "double2";
  }
}

bool jit_enabled() {
  char const* e = std::getenv("SPED_JIT");
  return !(e && e[0] == '0');
}

// NVRTC: generated program + matvec_kernel.cuh -> sm_100a cubin (empty on failure)
// On-disk cache of compiled modules: $SPED_CACHE_DIR, else $XDG_CACHE_HOME/sped-b200, else
// ~/.cache/sped-b200; the key hashes every source byte and option.  SPED_CACHE_DIR="" disables it.
u64 fnv1a(void const* data, size_t n, u64 h = 0xcbf29ce484222325ull) {
  auto p = static_cast<unsigned char const*>(data);
  for (size_t i = 0; i < n; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
  return h;
}

std::string cubin_cache_path(std::string const& header) {
  char const* dir = std::getenv("SPED_CACHE_DIR");
  std::string base;
  if (dir) {
    if (!*dir) return "";
    base = dir;
  } else if (char const* xdg = std::getenv("XDG_CACHE_HOME")) {
    base = std::string(xdg) + "/sped-b200";
  } else if (char const* home = std::getenv("HOME")) {
    base = std::string(home) + "/.cache/sped-b200";
  } else {
    return "";
  }
  // the key covers every source byte, every option and the compiler that would produce the module
  int major = 0, minor = 0;
  NvrtcApi& api = nvrtc();
  if (api.handle && api.Version) api.Version(&major, &minor);
  std::string const options = "sm_100a c++17 lineinfo v2 nvrtc " + std::to_string(major) + "." + std::to_string(minor);
  u64 h = fnv1a(header.data(), header.size());
  h = fnv1a(k_src_device_types, std::strlen(k_src_device_types), h ^ 0xff);
  h = fnv1a(k_src_matvec_kernel, std::strlen(k_src_matvec_kernel), h ^ 0xff);
  h = fnv1a(options.data(), options.size(), h ^ 0xff);
  std::error_code ec;  // no shell, no fork: the process may hold CUDA and NCCL state
  std::filesystem::create_directories(base, ec);
  if (ec) return "";
  char name[32];
  std::snprintf(name, sizeof name, "/%016llx.cubin", (unsigned long long)h);
  return base + name;
}

// cache file = "SPEDCUB2" | payload size | FNV-1a of the payload | payload; anything else is ignored
constexpr char kCacheMagic[8] = {'S', 'P', 'E', 'D', 'C', 'U', 'B', '2'};

std::vector<char> read_cached_cubin(std::string const& path) {
  std::vector<char> cubin;
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return cubin;
  char magic[8];
  u64 size = 0, sum = 0;
  bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, kCacheMagic, 8) == 0 && std::fread(&size, 8, 1, f) == 1 &&
            std::fread(&sum, 8, 1, f) == 1 && size > 0 && size < ((u64)1 << 31);
  if (ok) {
    cubin.resize((size_t)size);
    ok = std::fread(cubin.data(), 1, cubin.size(), f) == cubin.size() && std::fgetc(f) == EOF &&
         fnv1a(cubin.data(), cubin.size()) == sum;
  }
  std::fclose(f);
  if (!ok) {
    cubin.clear();
    SPED_LOG("jit: ignoring damaged or foreign cache entry %s", path.c_str());
  }
  return cubin;
}

void write_cached_cubin(std::string const& path, std::vector<char> const& cubin) {
  // published atomically: several ranks may compile the same module at the same time
  // (process AND thread: two bases with the same program may be compiling in this process at once)
  std::string tmp = path + "." + std::to_string((long)getpid()) + "." +
                    std::to_string((unsigned long long)std::hash<std::thread::id>()(std::this_thread::get_id())) + ".tmp";
  FILE* f = std::fopen(tmp.c_str(), "wb");
  if (!f) return;
  u64 const size = cubin.size(), sum = fnv1a(cubin.data(), cubin.size());
  bool ok = std::fwrite(kCacheMagic, 1, 8, f) == 8 && std::fwrite(&size, 8, 1, f) == 1 && std::fwrite(&sum, 8, 1, f) == 1 &&
            std::fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
  ok = (std::fclose(f) == 0) && ok;
  if (!ok || std::rename(tmp.c_str(), path.c_str()) != 0) std::remove(tmp.c_str());
}

// generated header of one module: kind, storage type, columns, and the canonicalisation as code
std::string module_header(Basis const& b, int kind, int dtype, int nb) {
  std::string program = Gen(b.program).run();
  return std::string("#define SPED_JIT_KIND ") + std::to_string(kind) + "\n#define SPED_T " + dtype_name(dtype) +
         "\n#define SPED_NB " + std::to_string(nb) + "\n" + program;
}

// header -> cubin (disk cache, else NVRTC).  Touches no basis and no CUDA context: safe on any thread
// once nvrtc() has been called on the thread that starts it.
std::vector<char> compile_header(std::string const& header) {
  std::vector<char> cubin;
  std::string cache_file = cubin_cache_path(header);
  if (!cache_file.empty()) {
    cubin = read_cached_cubin(cache_file);
    if (!cubin.empty()) {
      SPED_LOG("jit: reusing %s", cache_file.c_str());
      return cubin;
    }
  }
  NvrtcApi& api = nvrtc();
  if (!api.handle) {
    SPED_LOG("jit: NVRTC not available, using the interpreted-program kernel");
    return cubin;
  }
  std::string main_src =
      "#define SPED_JIT 1\n#include \"device_types.h\"\n#include \"sped_jit_program.h\"\n#include \"matvec_kernel.cuh\"\n";
  char const* hdr_src[] = {k_src_device_types, header.c_str(), k_src_matvec_kernel};
  char const* hdr_names[] = {"device_types.h", "sped_jit_program.h", "matvec_kernel.cuh"};
  nvrtcProgram prog = nullptr;
  if (api.CreateProgram(&prog, main_src.c_str(), "sped_matvec_jit.cu", 3, hdr_src, hdr_names) != NVRTC_SUCCESS) return cubin;
  char const* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device"};
  nvrtcResult rc = api.CompileProgram(prog, 4, opts);
  if (rc != NVRTC_SUCCESS || g_logging) {
    size_t n = 0;
    api.GetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) api.GetProgramLog(prog, &log[0]);
    if (rc != NVRTC_SUCCESS) {
      std::fprintf(stderr, "[sped] jit: NVRTC compilation failed, using the interpreted-program kernel\n%s\n", log.c_str());
      api.DestroyProgram(&prog);
      return cubin;
    }
  }
  size_t size = 0;
  api.GetCUBINSize(prog, &size);
  cubin.resize(size);
  api.GetCUBIN(prog, cubin.data());
  api.DestroyProgram(&prog);
  if (!cache_file.empty()) write_cached_cubin(cache_file, cubin);
  if (char const* dump = std::getenv("SPED_JIT_DUMP")) {  // inspection: cuobjdump -sass <file>
    if (FILE* f = std::fopen(dump, "wb")) {
      std::fwrite(cubin.data(), 1, cubin.size(), f);
      std::fclose(f);
    }
  }
  return cubin;
}

std::vector<char> compile_cubin(Basis& b, int kind, int dtype, int nb) { return compile_header(module_header(b, kind, dtype, nb)); }

// Compilations still running when the process ends are waited for before the statics they use go
// away (a basis handle may be leaked, or destroyed by a finaliser at exit).
struct InFlight {
  std::mutex m;
  std::vector<std::shared_future<std::vector<char>>> futures;
  void add(std::shared_future<std::vector<char>> f) {
    std::lock_guard<std::mutex> lock(m);
    futures.erase(std::remove_if(futures.begin(), futures.end(),
                                 [](auto const& x) { return x.wait_for(std::chrono::seconds(0)) == std::future_status::ready; }),
                  futures.end());
    futures.push_back(std::move(f));
  }
  ~InFlight() {
    for (auto& f : futures) f.wait();
  }
};
// One trivial synchronous compilation before the registry below is constructed: NVRTC creates its
// lazily initialised statics (and registers their destructors) now, so that at exit the registry --
// registered later, hence run earlier -- has waited for every compile thread before they go away.
void warm_up_nvrtc() {
  NvrtcApi& api = nvrtc();
  if (!api.handle) return;
  nvrtcProgram prog = nullptr;
  if (api.CreateProgram(&prog, "extern \"C\" __global__ void sped_warm_up() {}\n", "sped_warm_up.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) return;
  char const* opts[] = {"--gpu-architecture=sm_100a"};
  api.CompileProgram(prog, 1, opts);
  api.DestroyProgram(&prog);
}
InFlight& in_flight() {
  static bool const warmed = (warm_up_nvrtc(), true);  // also loads NVRTC: constructed first, destroyed last
  (void)warmed;
  static InFlight x;
  return x;
}

// SPED_JIT_WAIT=1: always wait for the specialised kernel (tests that must exercise it; benchmarks of
// the kernel itself).  Default: never stall a launch on NVRTC when the interpreted kernel gets the work
// done sooner.
bool jit_always_wait() {
  char const* e = std::getenv("SPED_JIT_WAIT");
  return e && e[0] == '1';
}

void start_compile(Basis& b, Entry& e, int kind, int dtype, int nb) {
  auto p = std::make_shared<Pending>();
  std::string header = module_header(b, kind, dtype, nb);
  p->started = std::chrono::steady_clock::now();
  p->expected_seconds = 0.25 + 0.004 * (double)b.program.steps.size();
  InFlight& reg = in_flight();
  p->cubin = std::async(std::launch::async, [header = std::move(header)]() { return compile_header(header); }).share();
  reg.add(p->cubin);
  e.pending = std::move(p);
}

// cubin -> loaded module; called on the thread that launches (the device context must be current)
void finish_compile(Basis& b, Entry& out, int kind, int dtype, int nb) {
  std::vector<char> cubin = out.pending->cubin.get();
  double const dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - out.pending->started).count();
  out.pending.reset();
  size_t size = cubin.size();
  if (cubin.empty()) {
    out.failed = true;
    return;
  }
  cudaError_t e = cudaLibraryLoadData(&out.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e == cudaSuccess) e = cudaLibraryGetKernel(&out.kernel, out.lib, kind == kKindFill ? "sped_cache_fill_jit" : "sped_matvec_jit");
  if (e != cudaSuccess) {
    std::fprintf(stderr, "[sped] jit: loading the specialised kernel failed (%s), using the interpreted-program kernel\n",
                 cudaGetErrorString(e));
    cudaGetLastError();
    out.failed = true;
    out.kernel = nullptr;
    return;
  }
  SPED_LOG("jit: specialised %s (%s, %d columns, %zu steps) ready %.2f s after it was requested, cubin %zu bytes",
           kind == kKindFill ? "cache fill" : "matvec", dtype_name(dtype), nb, b.program.steps.size(), dt, size);
}

// The specialised kernel, or null: not available (no NVRTC, SPED_JIT=0, trivial group, failure) -- or
// still compiling while the interpreted kernel is expected to finish the work at hand sooner.
// `images` = canonicalisation steps the launches at hand perform (rows x transitions per row x group
// order x traversals); the interpreted kernel needs about 0.55 ps more per image than the specialised
// one (6x6 on B200: 358 against 167 ms for 3.4e11 images).  images < 0: request only, never wait.
void* jit_kernel(Basis& b, int kind, int dtype, int nb, double images) {
  if (!jit_enabled() || b.trivial()) return nullptr;
  std::lock_guard<std::mutex> lock(g_jit_mutex);
  if (!b.jit_cache) b.jit_cache = std::make_shared<Cache>();
  Cache& c = *static_cast<Cache*>(b.jit_cache.get());
  auto key = std::make_tuple(kind, dtype, nb);
  auto it = c.entries.find(key);
  if (it == c.entries.end()) {
    it = c.entries.emplace(key, Entry{}).first;
    start_compile(b, it->second, kind, dtype, nb);
  }
  Entry& e = it->second;
  if (e.pending) {
    bool const ready = e.pending->cubin.wait_for(std::chrono::seconds(0)) == std::future_status::ready;
    if (!ready && !jit_always_wait()) {
      if (images < 0) return nullptr;
      double const elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - e.pending->started).count();
      double const remaining = std::max(0.05, e.pending->expected_seconds - elapsed);
      double const extra_if_interpreted = images * 0.55e-12;
      if (extra_if_interpreted < remaining) {
        SPED_LOG("jit: %s kernel still compiling (%.2f s so far); the interpreted kernel runs this time", kind == kKindFill ? "cache-fill" : "matvec", elapsed);
        return nullptr;
      }
    }
    finish_compile(b, e, kind, dtype, nb);
  }
  return e.failed ? nullptr : (void*)e.kernel;
}

}  // namespace

size_t jit_compile_only(Basis& b, int dtype, int nb) {
  std::vector<char> cubin = compile_cubin(b, kKindMatvec, dtype, nb);
  if (cubin.empty()) fail(SPED_INTERNAL_ERROR, "NVRTC compilation of the specialised kernel failed");
  std::vector<char> fill = compile_cubin(b, kKindFill, SPED_F64, 1);
  if (fill.empty()) fail(SPED_INTERNAL_ERROR, "NVRTC compilation of the specialised cache-fill kernel failed");
  return cubin.size();
}

// Source of the specialised canonicalisation (exposed for tests and inspection).
std::string jit_program_source(Basis const& b) { return Gen(b.program).run(); }

void* jit_matvec_kernel(Basis& b, int dtype, int nb, double images) { return jit_kernel(b, kKindMatvec, dtype, nb, images); }

// The cache fill does not depend on the storage type: one module per basis.
void* jit_cache_fill_kernel(Basis& b, double images) { return jit_kernel(b, kKindFill, SPED_F64, 1, images); }

// Called when a basis is built: the operator cache of any operator on it will want the fill kernel.
void jit_prefetch(Basis& b) {
  char const* force = std::getenv("SPED_JIT_PREFETCH");  // 0: never; 1: even without a device (host-side tests)
  if (force && force[0] == '0') return;
  int n_dev = 0;
  if (!(force && force[0] == '1') && (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)) {  // nothing could launch it
    cudaGetLastError();
    return;
  }
  jit_kernel(b, kKindFill, SPED_F64, 1, -1.0);
}

}  // namespace sped
