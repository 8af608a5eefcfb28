// caar_driver.cu — standalone B200 driver for compute_and_apply_rhs, the counterpart of the reference's
// cxx test driver (compute_and_apply_rhs_test/cxx/pointers_only/main.cpp:25-143): same command line flags,
// same closed-form synthetic data (data_structures.cpp:38-92,117-163), same printed norms
// (compute_and_apply_rhs.cpp:372-399) so the two outputs can be diffed — but the state lives on the GPUs.
//
//   caar_driver --tinman-num-elems=N --tinman-num-exec=M [--tinman-dump-res=yes|no]
//               [--caar-nlev=72] [--caar-mode=fast|strict] [--caar-gpus=G] [--caar-resident=yes|no] [--caar-checksums=yes|no]
//
// Multi-GPU: the element range is cut into G contiguous blocks [g*N/G, (g+1)*N/G) — the reference's own
// nets/nete partition hook (data_structures.hpp:58-66) — one host thread and one C-ABI handle per GPU, no
// communication while stepping. The only collective is one ncclAllReduce(sum) of the three squared norms
// over NVLink after the loop; rank 0 prints sqrt of the result.
#include <cuda_runtime.h>
#include <nccl.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "caar_b200.h"
#include "closed_form_init.hpp"

namespace {

using caar_host::HostData;
using caar_host::init_data;

struct Options {
  int num_elems = 10, num_exec = 1, nlev = 72, gpus = 1, mode = CAAR_MODE_FAST;
  bool dump = false, resident = true, checksums = false;
};

[[noreturn]] void die(const char* what, int rc) {
  std::fprintf(stderr, "caar_driver: %s failed (code %d): %s\n", what, rc, caar_last_error());
  std::exit(2);
}
#define OK(call)                       \
  do {                                 \
    if (int rc_ = (call)) die(#call, rc_); \
  } while (0)

struct Rank {
  caar_handle h = nullptr;
  int e0 = 0, n = 0, device = 0;
  caar_arrays host;
  double sumsq[3] = {0, 0, 0};
};

void print_norms(const double ss[3]) {
  std::printf("   ---> Norms:\n          ||v||_2  = %.17g\n          ||T||_2  = %.17g\n          ||dp||_2 = %.17g\n",
              std::sqrt(ss[0]), std::sqrt(ss[1]), std::sqrt(ss[2]));
}

// sum of the per-GPU squared norms: NCCL all-reduce over NVLink when G > 1
void reduce_norms(std::vector<Rank>& ranks, std::vector<ncclComm_t>& comms, int tl, double out[3]) {
  const int G = (int)ranks.size();
  for (auto& r : ranks) OK(caar_norms(r.h, tl, 0, r.n, r.sumsq));
  if (G == 1) {
    std::memcpy(out, ranks[0].sumsq, sizeof ranks[0].sumsq);
    return;
  }
  std::vector<double*> buf(G);
  std::vector<cudaStream_t> st(G);
  for (int g = 0; g < G; ++g) {
    cudaSetDevice(ranks[g].device);
    cudaMalloc(&buf[g], 3 * sizeof(double));
    cudaStreamCreate(&st[g]);
    cudaMemcpyAsync(buf[g], ranks[g].sumsq, 3 * sizeof(double), cudaMemcpyHostToDevice, st[g]);
  }
  ncclGroupStart();
  for (int g = 0; g < G; ++g) ncclAllReduce(buf[g], buf[g], 3, ncclDouble, ncclSum, comms[g], st[g]);
  ncclGroupEnd();
  for (int g = 0; g < G; ++g) {
    cudaSetDevice(ranks[g].device);
    cudaStreamSynchronize(st[g]);
  }
  cudaSetDevice(ranks[0].device);
  cudaMemcpy(out, buf[0], 3 * sizeof(double), cudaMemcpyDeviceToHost);
  for (int g = 0; g < G; ++g) {
    cudaSetDevice(ranks[g].device);
    cudaFree(buf[g]);
    cudaStreamDestroy(st[g]);
  }
}

// The other collective of the job (north star: "final allreduce of field checksums and energy norms"): per-GPU
// caar_checksums summed over ranks — sums, sums of squares and the two energy norms as doubles, the exact bit-pattern
// sums as 64-bit integers (wrap-around addition is associative, so the result does not depend on the partition).
void reduce_checksums(std::vector<Rank>& ranks, std::vector<ncclComm_t>& comms, int tl, caar_checksum* out) {
  const int G = (int)ranks.size();
  std::vector<caar_checksum> cs(G);
  for (int g = 0; g < G; ++g) OK(caar_checksums(ranks[g].h, tl, 0, ranks[g].n, &cs[g]));
  *out = cs[0];
  if (G == 1) return;
  std::vector<double*> fbuf(G);
  std::vector<unsigned long long*> ibuf(G);
  std::vector<cudaStream_t> st(G);
  for (int g = 0; g < G; ++g) {
    double f[16];
    std::memcpy(f, cs[g].sum, 7 * sizeof(double));
    std::memcpy(f + 7, cs[g].sumsq, 7 * sizeof(double));
    std::memcpy(f + 14, cs[g].energy, 2 * sizeof(double));
    cudaSetDevice(ranks[g].device);
    cudaMalloc(&fbuf[g], sizeof f);
    cudaMalloc(&ibuf[g], sizeof cs[g].bits);
    cudaStreamCreate(&st[g]);
    cudaMemcpy(fbuf[g], f, sizeof f, cudaMemcpyHostToDevice);
    cudaMemcpy(ibuf[g], cs[g].bits, sizeof cs[g].bits, cudaMemcpyHostToDevice);
  }
  ncclGroupStart();
  for (int g = 0; g < G; ++g) {
    ncclAllReduce(fbuf[g], fbuf[g], 16, ncclDouble, ncclSum, comms[g], st[g]);
    ncclAllReduce(ibuf[g], ibuf[g], 7, ncclUint64, ncclSum, comms[g], st[g]);
  }
  ncclGroupEnd();
  for (int g = 0; g < G; ++g) {
    cudaSetDevice(ranks[g].device);
    cudaStreamSynchronize(st[g]);
  }
  double f[16];
  cudaSetDevice(ranks[0].device);
  cudaMemcpy(f, fbuf[0], sizeof f, cudaMemcpyDeviceToHost);
  cudaMemcpy(out->bits, ibuf[0], sizeof out->bits, cudaMemcpyDeviceToHost);
  std::memcpy(out->sum, f, 7 * sizeof(double));
  std::memcpy(out->sumsq, f + 7, 7 * sizeof(double));
  std::memcpy(out->energy, f + 14, 2 * sizeof(double));
  for (int g = 0; g < G; ++g) {
    cudaSetDevice(ranks[g].device);
    cudaFree(fbuf[g]);
    cudaFree(ibuf[g]);
    cudaStreamDestroy(st[g]);
  }
}

void print_checksums(const caar_checksum& c) {
  static const char* const name[7] = {"dp3d", "v", "T", "eta_dot_dpdn", "omega_p", "phi", "vn0"};
  std::printf("   ---> Checksums (sum, sum of squares, bit-pattern sum mod 2^64):\n");
  for (int f = 0; f < 7; ++f)
    std::printf("          %-13s %.17g %.17g %016llx\n", name[f], c.sum[f], c.sumsq[f], c.bits[f]);
  std::printf("          energy: kinetic %.17g internal %.17g\n", c.energy[0], c.energy[1]);
}

void dump(const HostData& d) {
  static const char* const names[4] = {"elem_state_vx.txt", "elem_state_vy.txt", "elem_state_t.txt", "elem_state_dp3d.txt"};
  std::ofstream out[4];
  for (int f = 0; f < 4; ++f) {
    out[f].open(names[f]);
    if (!out[f].is_open()) {
      std::printf("Error! Cannot open '%s'.\n", names[f]);
      std::abort();
    }
    out[f].precision(6);
  }
  const int L = d.dims.nlev, tl = d.ctl.np1;
  for (int ie = 0; ie < d.dims.nelem; ++ie)
    for (int k = 0; k < L; ++k) {
      for (auto& o : out) o << "[" << ie << ", " << k << "]\n";
      for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) {
          const size_t n = (((size_t)ie * 3 + tl) * L + k) * 16 + i * 4 + j;
          out[0] << " " << d.f[7][2 * n];
          out[1] << " " << d.f[7][2 * n + 1];
          out[2] << " " << d.f[8][n];
          out[3] << " " << d.f[6][n];
        }
        for (auto& o : out) o << "\n";
      }
    }
}

bool all_digits(const char* s) {
  if (!*s) return false;
  for (; *s; ++s)
    if (*s < '0' || *s > '9') return false;
  return true;
}

}  // namespace

int main(int argc, char** argv) {
  Options o;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    const size_t eq = a.find('=');
    const std::string key = a.substr(0, eq), val = eq == std::string::npos ? "" : a.substr(eq + 1);
    auto yesno = [&](bool& dst) {
      if (val == "yes" || val == "YES") dst = true;
      else if (val == "no" || val == "NO") dst = false;
      else {
        std::printf(" ERROR! Unrecognized command line option '%s'.\n        Run with '--tinman-help' to see the available options.\n", a.c_str());
        std::exit(1);
      }
    };
    if (key == "--tinman-num-elems") {
      if (!all_digits(val.c_str())) {
        std::fprintf(stderr, "Expecting an unsigned integer after '--tinman-num-elems='.\n");
        return 1;
      }
      o.num_elems = std::atoi(val.c_str());
    } else if (key == "--tinman-num-exec") o.num_exec = std::atoi(val.c_str());
    else if (key == "--tinman-dump-res") yesno(o.dump);
    else if (key == "--caar-resident") yesno(o.resident);
    else if (key == "--caar-checksums") yesno(o.checksums);
    else if (key == "--caar-nlev") o.nlev = std::atoi(val.c_str());
    else if (key == "--caar-gpus") o.gpus = std::atoi(val.c_str());
    else if (key == "--caar-mode") o.mode = (val == "strict") ? CAAR_MODE_STRICT : CAAR_MODE_FAST;
    else if (key == "--tinman-help") {
      std::printf("  --tinman-num-elems=N  : the number of elements (default=10)\n"
                  "  --tinman-dump-res=val : whether to dump results to file (default=no)\n"
                  "  --tinman-num-exec=N   : number of times to execute (default=1)\n"
                  "  --caar-nlev=L         : vertical levels (default=72)\n"
                  "  --caar-mode=fast|strict, --caar-gpus=G, --caar-resident=yes|no, --caar-checksums=yes|no\n"
                  "  --tinman-help         : prints this message\n");
      return 0;
    }
  }
  if (o.num_elems < 1) {
    std::fprintf(stderr, "Invalid number of elements: %d\n", o.num_elems);
    return 1;
  }
  const int ndev = caar_device_count();
  if (ndev < 1) {
    std::fprintf(stderr, "caar_driver: no CUDA device (%s); there is no CPU path\n", caar_last_error());
    return 2;
  }
  if (o.gpus < 1 || o.gpus > ndev || o.gpus > o.num_elems) {
    std::fprintf(stderr, "caar_driver: --caar-gpus=%d not possible (%d devices, %d elements)\n", o.gpus, ndev, o.num_elems);
    return 1;
  }

  std::printf(" --- Initializing data...\n");
  HostData d;
  init_data(d, o.num_elems, o.nlev);
  const int G = o.gpus;
  std::vector<Rank> ranks(G);
  std::vector<ncclComm_t> comms(G);
  if (G > 1) {
    std::vector<int> devs(G);
    for (int g = 0; g < G; ++g) devs[g] = g;
    if (ncclCommInitAll(comms.data(), G, devs.data()) != ncclSuccess) {
      std::fprintf(stderr, "caar_driver: ncclCommInitAll failed\n");
      return 2;
    }
  }
  for (int g = 0; g < G; ++g) {
    Rank& r = ranks[g];
    r.device = g;
    r.e0 = (int)((long long)o.num_elems * g / G);
    r.n = (int)((long long)o.num_elems * (g + 1) / G) - r.e0;
    r.host = d.slice(r.e0);
    caar_dims dims = d.dims;
    dims.nelem = r.n;
    OK(caar_create(&r.h, &dims, g));
    OK(caar_set_params(r.h, &d.c, d.dvv, d.ps0, d.hyai.data()));
    OK(caar_upload(r.h, &r.host, CAAR_F_ALL));
  }
  if (!o.resident)  // host arrays travel every call: page-lock them in place so the copies are DMA at PCIe speed
    for (int i = 0; i < CAAR_NUM_FIELDS; ++i)
      if (caar_host_register(d.f[i].data(), d.f[i].size() * sizeof(double)) != CAAR_OK)
        std::fprintf(stderr, "caar_driver: could not page-lock array %d (%s); continuing with pageable memory\n", i,
                     caar_last_error());
  double ss[3];
  reduce_norms(ranks, comms, d.ctl.np1, ss);
  print_norms(ss);

  std::printf(" --- Performing computations... (%d executions of the main loop on %d elements)\n", o.num_exec, o.num_elems);
  const auto t0 = std::chrono::steady_clock::now();
  {
    std::vector<std::thread> pool;
    for (int g = 0; g < G; ++g)
      pool.emplace_back([&, g]() {
        Rank& r = ranks[g];
        caar_control ctl = d.ctl;
        ctl.nets = 0;
        ctl.nete = r.n;
        if (o.resident) {
          OK(caar_run(r.h, &ctl, o.num_exec, o.mode));
          OK(caar_sync(r.h));
        } else {
          for (int it = 0; it < o.num_exec; ++it) OK(caar_run_host(r.h, &r.host, &ctl, o.mode, 0));
        }
      });
    for (auto& t : pool) t.join();
  }
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf("   ---> compute_and_apply_rhs execution total time: %g s (wall, %d GPU%s, %.4g elem*lev updates/s)\n", sec, G,
              G > 1 ? "s" : "", (double)o.num_elems * o.nlev * o.num_exec / sec);
  reduce_norms(ranks, comms, d.ctl.np1, ss);
  print_norms(ss);
  if (o.checksums) {
    caar_checksum cs;
    reduce_checksums(ranks, comms, d.ctl.np1, &cs);
    print_checksums(cs);
  }

  if (o.dump) {
    std::printf(" --- Dumping results to file...\n");
    for (auto& r : ranks) OK(caar_download(r.h, &r.host, CAAR_F_MUTATED));
    dump(d);
  }
  std::printf(" --- Cleaning up data...\n");
  if (!o.resident)
    for (int i = 0; i < CAAR_NUM_FIELDS; ++i) caar_host_unregister(d.f[i].data());
  for (auto& r : ranks) caar_destroy(r.h);
  if (G > 1)
    for (auto& c : comms) ncclCommDestroy(c);
  return 0;
}
