// LBMCL<T>: the simulation driver of the LBMCL host program, re-implemented over the C ABI of
// liblbm_b200.so (include/lbm_b200.h) instead of the OpenCL C++ bindings.
//
// Public surface = what reference main.cpp drives (reference lbmcl.hpp): the 14-argument constructor
// (:338-388), setupSimulation (:392), printConfiguration (:623), performSimulation (:485-521),
// waitCompletion (:525), performSimulationAndWait (:538), totalTimeMS (:548), kernelsTimeMS (:562),
// kernelsTimingsMS (:580), MLUPS (:604), kernelsMLUPS (:617), statistics (:648).  File formats are the
// reference's: lbmcl.<it>.vti (:261-334), map.dump (:159-203), f_<it>.dump (:206-258).
//
// Differences, all forced by the change of device runtime:
//   * no platform; the device id is a CUDA ordinal and a negative id means device 0 instead of an
//     interactive prompt (reference CLUtil.hpp:126-138);
//   * the work-group size is a hint (lbm_block_shape reports the CUDA block actually used);
//   * optional z-slab decomposition over several GPUs (`gpus` > 1);
//   * device errors are reported as "file:line what(code) - cudaErrorName" by the library and end the
//     program with a non-zero status here (reference CLUtil.hpp:101-117 does the same with CL names);
//   * output is pipelined (the reference blocks in storeData, lbmcl.hpp:261-334): rho/u are
//     read back asynchronously into one of two page-locked buffers on a copy stream while the device already
//     computes the next `every` iterations, and a writer thread formats and writes the VTI file meanwhile;
//     the ASCII body is formatted by several threads with an exact "%.16e" routine (fmt_e16.hpp) -- the bytes
//     are the reference's, the wall time is not.
#pragma once

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/lbm_b200.h"
#include "vti_writer.hpp"

#define LBM_INITIALIZE_KERNEL_NAME "initialize"
#define LBM_COMPUTE_KERNEL_NAME "compute"

template <typename T>
class LBMCL {
    static_assert(std::is_same<T, float>::value || std::is_same<T, double>::value,
                  "Only float or double data type is valid.");

    size_t dim;
    T viscosity;
    T velocity;
    size_t iterations;
    size_t every;
    std::string vtk_path;
    size_t lws[3];
    size_t stride;
    bool optimize;
    std::string dump_path;
    bool dump_map;
    bool dump_f;
    bool dump_data;
    int gpus;
    bool aa;

    lbm_ctx *ctx = nullptr;
    lbm_group *group = nullptr;
    std::string device_name = "?";
    int64_t device_bytes = 0;

    std::vector<T> f_values;
    std::vector<int32_t> map_values;

    // snapshot pipeline: two host buffers (page-locked when possible), one writer thread at a time
    T *rho_buf[2] = {nullptr, nullptr};
    T *u_buf[2] = {nullptr, nullptr};
    bool buf_pinned = false;
    std::thread writer;
    size_t n_snapshots = 0;

    static constexpr size_t Q = 19, D = 3;

    size_t cells() const { return dim * dim * dim; }
    size_t wet_dim() const { return (dim - 2) * (dim - 2) * (dim - 2); }
    static size_t digits(size_t v) { return v > 0 ? (size_t)std::log10((double)v) + 1 : 1; }  // lbmcl.hpp:13
    static bool is_pow2(size_t x) { return x && !(x & (x - 1)); }
    static size_t floor_pow2(size_t x)
    {
        size_t p = 1;
        while (p * 2 <= x && p * 2 != 0) p *= 2;
        return x ? p : 0;
    }
    static size_t log2i(size_t x)
    {
        size_t n = 0;
        for (; x > 1; x >>= 1) ++n;
        return n;
    }

    [[noreturn]] void die(int code, const char *what) const
    {
        const char *msg = group ? lbm_group_last_error(group) : lbm_last_error(ctx);
        std::cerr << what << " failed (" << code << "): " << (msg ? msg : "") << std::endl;
        std::exit(code == 0 ? 1 : (code < 0 ? -code : code));
    }
    void check(int rc, const char *what) const
    {
        if (rc != LBM_OK) die(rc, what);
    }

    // what lbmcl.hpp:131-156 hands to the OpenCL compiler; kept as the description of the
    // specialisation (the same values parameterise the CUDA kernels)
    std::string kernelOptionsStr() const
    {
        std::stringstream o;
        o << "-DDIM=" << dim << " -DLWS=" << lws[0] << " -DSTRIDE_DIV=" << log2i(stride) << " -DSTRIDE_MOD="
          << (stride - 1) << " -DVISCOSITY=" << viscosity << " -DVELOCITY=" << velocity << " ";
        o << (std::is_same<T, float>::value ? "-DFP_SINGLE " : "-DFP_DOUBLE ");
        if (optimize) o << "-use_fast_math ";
        o << "[sm_100a";
        if (gpus > 1) o << ", " << gpus << " z-slabs";
        if (aa) o << ", in-place AA";
        o << "]";
        return o.str();
    }

    std::string numbered(const std::string &dir, const char *stem, size_t it, const char *ext) const
    {
        std::stringstream s;
        s << dir << "/" << stem << std::setw((int)digits(iterations)) << std::setfill('0') << it << ext;
        return s.str();
    }

    int n_ctx() const { return group ? lbm_group_size(group) : 1; }
    lbm_ctx *ctx_at(int i) const { return group ? lbm_group_ctx(group, i) : ctx; }

    // map.dump, lbmcl.hpp:159-203.  A missing directory is a silent no-op there (no check on the
    // ofstream) and here.
    void storeMap()
    {
        check(lbm_read_map(group ? lbm_group_ctx(group, 0) : ctx, map_values.data()), "read_map");
        FILE *fp = std::fopen((dump_path + "/map.dump").c_str(), "w");
        if (!fp) return;
        std::fputs("# FLUID       1\n# MOVING      2\n# BOUNDARY    3\n# WALL        4\n# CORNER      5\n\n", fp);
        std::string line;
        for (size_t z = 0; z < dim; ++z) {
            for (size_t y = 0; y < dim; ++y) {
                line.clear();
                for (size_t x = 0; x < dim; ++x) {
                    const int t = map_values[x + y * dim + z * dim * dim];
                    // the reference assigns the categories in sequence, later ones win (:188-193):
                    // lid cells carry the FRONT bit and therefore print 3, not 2
                    int v = 0;
                    if (t == 0x1) v = 1;
                    if (t & 0x2) v = 2;
                    if (t & 0x3f0) v = 3;
                    if (t == 0x8) v = 4;
                    if (t == 0x4) v = 5;
                    line += (char)('0' + v);
                    line += ' ';
                }
                line += '\n';
                std::fputs(line.c_str(), fp);
            }
            std::fputc('\n', fp);
        }
        std::fputc('\n', fp);
        std::fclose(fp);
    }

    // f_<it>.dump, lbmcl.hpp:206-258: the populations iteration `it` reads, CSoA order.
    void storeF(size_t iteration)
    {
        if (group) check(lbm_group_read_f(group, f_values.data()), "read_f");
        else check(lbm_read_f(ctx, f_values.data()), "read_f");
        FILE *fp = std::fopen(numbered(dump_path, "f_", iteration, ".dump").c_str(), "w");
        if (!fp) return;
        const int dd = (int)digits(dim);
        const size_t coord_spaces = (size_t)dd * 3 + 5;
        std::string head(coord_spaces, ' ');
        char buf[128];
        for (size_t q = 0; q < Q; ++q) {
            std::snprintf(buf, sizeof buf, "%8zu ", q);
            head += buf;
        }
        head += '\n';
        std::string line;
        for (size_t z = 0; z < dim; ++z) {
            for (size_t y = 0; y < dim; ++y) {
                std::fputs(head.c_str(), fp);
                for (size_t x = 0; x < dim; ++x) {
                    const size_t id = x + y * dim + z * dim * dim;
                    std::snprintf(buf, sizeof buf, "%*s%u,%u,%u) ", dd, "(", (unsigned)x, (unsigned)y, (unsigned)z);
                    line = buf;
                    for (size_t q = 0; q < Q; ++q) {
                        const T v = f_values[((id / stride) * Q + q) * stride + (id & (stride - 1))];
                        std::snprintf(buf, sizeof buf, "%8.6f ", (double)v);
                        line += buf;
                    }
                    line += '\n';
                    std::fputs(line.c_str(), fp);
                }
                std::fputc('\n', fp);
            }
            std::fputc('\n', fp);
        }
        std::fputc('\n', fp);
        std::fclose(fp);
    }

    // ---- VTI output, lbmcl.hpp:261-334 (vti_writer.hpp) ----
    void writeVTI(size_t iteration, const T *rho, const T *u) const
    {
        lbm_vti::write_vti<T>(numbered(vtk_path, "lbmcl.", iteration, ".vti"), dim, rho, u, 0);
    }

    // The reference's storeData(it) (blocking reads + file, lbmcl.hpp:261-334) as a pipeline stage.  Called
    // right after the iterations up to `iteration` have been enqueued; `enqueue_more` enqueues the work that may
    // overlap with this snapshot's read-back and file (the next batch of iterations), or is empty.
    template <typename F>
    void storeData(size_t iteration, F enqueue_more)
    {
        const size_t slot = n_snapshots & 1;  // its previous user (snapshot n-2) was joined before snapshot n-1 started
        for (int i = 0; i < n_ctx(); ++i)     // every slab copies its own planes into the global arrays
            check(lbm_read_macros_async(ctx_at(i), rho_buf[slot], u_buf[slot]), "read_rho/read_u");
        enqueue_more();                        // the device carries on while the copy runs
        for (int i = 0; i < n_ctx(); ++i) check(lbm_read_wait(ctx_at(i)), "read_rho/read_u");
        if (writer.joinable()) writer.join();  // one file at a time, in order
        const T *r = rho_buf[slot], *v = u_buf[slot];
        writer = std::thread([this, iteration, r, v] { writeVTI(iteration, r, v); });
        ++n_snapshots;
    }
    void storeData(size_t iteration)
    {
        storeData(iteration, [] {});
    }

public:
    LBMCL(size_t dim, T viscosity, T velocity, size_t iterations, size_t every, std::string vtk_path = "",
          size_t lwx = 1, size_t lwy = 1, size_t lwz = 1, size_t stride = 32, bool optimize = true,
          std::string dump_path = "", bool dump_map = false, bool dump_f = false, int gpus = 1, bool aa = false)
        : dim(dim), viscosity(viscosity), velocity(velocity), iterations(iterations), every(every),
          vtk_path(std::move(vtk_path)), stride(stride), optimize(optimize), dump_path(std::move(dump_path)),
          dump_map(dump_map), dump_f(dump_f), dump_data(every != 0), gpus(gpus < 1 ? 1 : gpus), aa(aa)
    {
        if (!is_pow2(this->dim)) {  // lbmcl.hpp:364-367
            this->dim = floor_pow2(this->dim);
            std::cout << "dim is rounded to the previous power of 2: " << this->dim << std::endl;
        }
        lws[0] = lwx ? lwx : 1;
        lws[1] = lwy ? lwy : 1;
        lws[2] = lwz ? lwz : 1;
        if (lws[0] * lws[1] * lws[2] > this->dim * this->dim * this->dim) {  // lbmcl.hpp:375-378
            std::cerr << "Please enter a good work_group_size to run the simulation" << std::endl;
            std::exit(-1);
        }
        if (!is_pow2(this->stride)) {  // lbmcl.hpp:384-387
            this->stride = floor_pow2(this->stride);
            std::cout << "stride is rounded to the previous power of 2: " << this->stride << std::endl;
        }
    }

    LBMCL(const LBMCL &) = delete;
    LBMCL &operator=(const LBMCL &) = delete;

    ~LBMCL()
    {
        if (writer.joinable()) writer.join();
        for (int i = 0; i < 2; ++i) {
            if (buf_pinned) {
                lbm_host_free(rho_buf[i]);
                lbm_host_free(u_buf[i]);
            } else {
                delete[] rho_buf[i];
                delete[] u_buf[i];
            }
        }
        if (group) lbm_group_destroy(group);
        if (ctx) lbm_destroy(ctx);
    }

    // lbmcl.hpp:392-482.  platformID is accepted for compatibility and ignored.
    void setupSimulation(int /*platformID*/, int deviceID)
    {
        lbm_params p;
        lbm_default_params(&p);
        p.dim = (int32_t)dim;
        p.precision = std::is_same<T, float>::value ? LBM_F32 : LBM_F64;
        p.fast_math = optimize ? 1 : 0;
        p.viscosity = (double)viscosity;
        p.velocity = (double)velocity;
        p.stride = (int64_t)stride;
        p.block_x = (int32_t)lws[0];
        p.block_y = (int32_t)lws[1];
        p.block_z = (int32_t)lws[2];
        p.device = deviceID < 0 ? 0 : deviceID;
        p.variant = aa ? LBM_VARIANT_AA : LBM_VARIANT_AUTO;
        if (const char *v = std::getenv("LBM_VARIANT")) p.variant = std::atoi(v);
        // -w is a hint by default; LBM_EXACT_BLOCK=1 makes the library honour it (block-shape sweeps)
        if (const char *v = std::getenv("LBM_EXACT_BLOCK")) p.reserved[1] = std::atoi(v) == 1 ? 1 : 0;
        char name[256] = "?";
        if (gpus > 1) {
            std::vector<int32_t> devs;
            for (int i = 0; i < gpus; ++i) devs.push_back(p.device + i);
            const int rc = lbm_group_create(&p, devs.data(), gpus, &group);
            if (rc != LBM_OK) {
                std::cerr << "lbm_group_create failed (" << rc << "): " << lbm_group_last_error(nullptr) << std::endl;
                std::exit(-rc);
            }
            lbm_device_name(lbm_group_ctx(group, 0), name, sizeof name);
            for (int i = 0; i < gpus; ++i) device_bytes += lbm_device_bytes(lbm_group_ctx(group, i));
        } else {
            const int rc = lbm_create(&p, &ctx);
            if (rc != LBM_OK) {
                std::cerr << "lbm_create failed (" << rc << "): " << lbm_last_error(nullptr) << std::endl;
                std::exit(-rc);
            }
            lbm_device_name(ctx, name, sizeof name);
            device_bytes = lbm_device_bytes(ctx);
        }
        device_name = name;
        if (dump_map) map_values.resize(cells());
        if (dump_f) f_values.resize(cells() * Q);
        if (dump_data) {
            // two snapshot buffers, page-locked so that the read-back is a real asynchronous DMA
            void *p[4] = {nullptr, nullptr, nullptr, nullptr};
            bool ok = true;
            for (int i = 0; i < 4 && ok; ++i)
                ok = lbm_host_alloc((i < 2 ? cells() : cells() * D) * sizeof(T), &p[i]) == LBM_OK;
            if (ok) {
                buf_pinned = true;
                for (int i = 0; i < 2; ++i) {
                    rho_buf[i] = static_cast<T *>(p[i]);
                    u_buf[i] = static_cast<T *>(p[2 + i]);
                }
            } else {  // pageable memory works too (the copies then block a little longer)
                for (int i = 0; i < 4; ++i) lbm_host_free(p[i]);
                for (int i = 0; i < 2; ++i) {
                    rho_buf[i] = new T[cells()];
                    u_buf[i] = new T[cells() * D];
                }
            }
        }
    }

    // lbmcl.hpp:485-521.  Kernel launches are asynchronous.  storeMap / storeF block like the reference's blocking
    // enqueueReadBuffer calls do; storeData is pipelined: while snapshot k is copied, formatted and written, the
    // device computes the iterations up to snapshot k+1.
    void performSimulation()
    {
        if (group) check(lbm_group_init(group), LBM_INITIALIZE_KERNEL_NAME);
        else check(lbm_init(ctx), LBM_INITIALIZE_KERNEL_NAME);
        if (dump_map) storeMap();
        if (dump_f) {
            // -f: one launch per iteration, the lattice is fetched before each (lbmcl.hpp:503, 517-519)
            if (dump_data) storeData(0);
            storeF(0);
            for (size_t it = 1; it <= iterations; ++it) {
                storeF(it);  // f_<it>.dump holds what iteration `it` reads
                const int flag = (dump_data && it % every == 0) ? 1 : 0;
                if (group) check(lbm_group_run(group, 1, flag), LBM_COMPUTE_KERNEL_NAME);
                else check(lbm_step(ctx, flag), LBM_COMPUTE_KERNEL_NAME);
                if (flag) storeData(it);
            }
            return;
        }
        // batches of iterations up to the next one that stores data, each enqueued as soon as the previous
        // snapshot's copy has been queued
        size_t it = 0;
        auto run_batch = [&]() {
            if (it >= iterations) return;
            size_t chunk = iterations - it;
            if (dump_data) {
                const size_t to_next = every - (it % every);
                if (to_next < chunk) chunk = to_next;
            }
            const int ev = dump_data ? (int)every : 0;
            if (group) check(lbm_group_run(group, (int)chunk, ev), LBM_COMPUTE_KERNEL_NAME);
            else check(lbm_run(ctx, (int)chunk, ev), LBM_COMPUTE_KERNEL_NAME);
            it += chunk;
        };
        if (dump_data) {
            storeData(0, run_batch);  // lbmcl.hpp:502; the first batch is enqueued behind the copy
            while (it > 0 && it % every == 0) {
                const size_t at = it;
                storeData(at, run_batch);  // lbmcl.hpp:513-515
                if (at >= iterations) break;
            }
        } else {
            run_batch();
        }
    }

    void waitCompletion()  // lbmcl.hpp:525-532
    {
        if (writer.joinable()) writer.join();
        if (group) check(lbm_group_sync(group), "finish");
        else check(lbm_sync(ctx), "finish");
        // "Total time" spans the output work like the reference's does (lbmcl.hpp:548-556)
        if (dump_data)
            for (int i = 0; i < n_ctx(); ++i) check(lbm_mark_end(ctx_at(i)), "finish");
    }

    void performSimulationAndWait()
    {
        performSimulation();
        waitCompletion();
    }

    // first event start -> last event end, read-backs and file writing in between included
    // (lbmcl.hpp:548-556)
    double totalTimeMS()
    {
        double total = 0.0;
        if (group) check(lbm_group_time_ms(group, &total, nullptr), "time");
        else check(lbm_time_ms(ctx, &total, nullptr), "time");
        return total;
    }

    // sum over the compute launches only (lbmcl.hpp:562-574)
    double kernelsTimeMS()
    {
        double k = 0.0;
        if (group) check(lbm_group_time_ms(group, nullptr, &k), "time");
        else check(lbm_time_ms(ctx, nullptr, &k), "time");
        return k;
    }

    // lbmcl.hpp:580-593: one (name, milliseconds) entry per timed enqueue.  With -f every iteration is
    // its own launch; otherwise launches are enqueued and timed in asynchronous batches (one entry each).
    std::vector<std::pair<std::string, double>> kernelsTimingsMS()
    {
        std::vector<std::pair<std::string, double>> t;
        if (group) {
            t.emplace_back(LBM_COMPUTE_KERNEL_NAME, kernelsTimeMS());
            return t;
        }
        int64_t n = 0;
        check(lbm_launch_times_ms(ctx, nullptr, 0, &n), "time");
        std::vector<double> ms((size_t)n);
        if (n > 0) check(lbm_launch_times_ms(ctx, ms.data(), n, &n), "time");
        for (double v : ms) t.emplace_back(LBM_COMPUTE_KERNEL_NAME, v);
        return t;
    }

    double MLUPS() { return (wet_dim() * iterations) / (totalTimeMS() * 1000); }           // lbmcl.hpp:604-607
    double kernelsMLUPS() { return (wet_dim() * iterations) / (kernelsTimeMS() * 1000); }  // lbmcl.hpp:617-620

    void printConfiguration()  // lbmcl.hpp:623-645
    {
        const char *prec = std::is_same<T, float>::value ? "single" : "double";
        int32_t blk[3] = {0, 0, 0}, vec = 0;
        lbm_block_shape(group ? lbm_group_ctx(group, 0) : ctx, blk, &vec);
        std::cout << std::boolalpha
                  << "kernel options   = " << kernelOptionsStr() << "\n"
                  << "device           = " << device_name << "\n"
                  << "dim              = " << dim << "\n"
                  << "viscosity        = " << viscosity << "\n"
                  << "velocity         = " << velocity << "\n"
                  << "Device Mem. (B)  = " << device_bytes << "\n"
                  << "Device Mem. (KB) = " << device_bytes / (1 << 10) << "\n"
                  << "Device Mem. (MB) = " << device_bytes / (1 << 20) << "\n"
                  << "iterations       = " << iterations << "\n"
                  << "work_group_size  = (" << lws[0] << ", " << lws[1] << ", " << lws[2] << ")\n"
                  << "stride           = " << stride << "\n"
                  << "precision        = " << prec << "\n"
                  << "optimize         = " << optimize << "\n"
                  << "every            = " << every << "\n"
                  << "VTK PATH         = " << vtk_path << "\n"
                  << "DUMP F           = " << dump_f << "\n"
                  << "DUMP MAP         = " << dump_map << "\n"
                  << "CUDA block       = (" << blk[0] << ", " << blk[1] << ", " << blk[2] << ") x " << vec
                  << " cell(s)/thread, " << gpus << " GPU(s)\n";
    }

    // lbmcl.hpp:648-669; the line benchmark.sh:111-176 aggregates
    std::string statistics(char separator)
    {
        const char *prec = std::is_same<T, float>::value ? "single" : "double";
        std::stringstream s;
        s << device_name << separator << prec << separator << dim << separator << iterations << separator << every
          << separator << std::setw(3) << std::setfill('0') << lws[0] << "," << std::setw(3) << std::setfill('0')
          << lws[1] << "," << std::setw(3) << std::setfill('0') << lws[2] << separator << stride << separator
          << optimize << separator << totalTimeMS() << separator << kernelsTimeMS() << separator << MLUPS()
          << separator << kernelsMLUPS() << "\n";
        return s.str();
    }
};
