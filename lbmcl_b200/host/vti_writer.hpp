// The VTI writer of the LBMCL host program: lbmcl.<it>.vti as the reference writes it (lbmcl.hpp:261-334) --
// ASCII ImageData, WholeExtent 0..DIM-3 per axis, `rho` then the 3-component `v` over the wet cube
// x,y,z in [1, DIM-2], x fastest, every value printed as std::scientific << std::setprecision(16) followed by a
// blank, one text line per x-row -- byte for byte, but written by several threads:
//   each worker formats one z-plane per round (exact "%.16e", fmt_e16.hpp), learns its file offset from the
//   sizes of the planes before it, and copies its text itself into a shared mapping of the file (the file is
//   first extended to an upper bound of its size and cut to the real size at the end), so that formatting AND
//   the copy into the page cache run in parallel -- write()/pwrite() calls on one file serialise on its inode
//   lock -- and the text never exists in memory as a whole.  Where the file cannot be mapped the workers fall
//   back to pwrite.
// ASCII output dominates every `-e N` run (1.5 GB per file at 256^3).
#pragma once

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/statvfs.h>
#include <unistd.h>

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "fmt_e16.hpp"

namespace lbm_vti {

struct Barrier {  // C++11 has none
    std::mutex m;
    std::condition_variable cv;
    unsigned n, count = 0, generation = 0;
    explicit Barrier(unsigned n) : n(n) {}
    void wait()
    {
        std::unique_lock<std::mutex> l(m);
        const unsigned g = generation;
        if (++count == n) {
            count = 0;
            ++generation;
            cv.notify_all();
        } else {
            cv.wait(l, [&] { return g != generation; });
        }
    }
};

// where the bytes of the file go: a shared mapping (map != nullptr) or pwrite on fd
struct Sink {
    int fd = -1;
    char *map = nullptr;
    size_t map_size = 0;
};

inline void write_at(const Sink &k, const char *p, size_t n, size_t off)
{
    if (k.map) {
        if (off + n <= k.map_size) std::memcpy(k.map + off, p, n);
        return;
    }
    const int fd = k.fd;
    while (n > 0) {
        const ssize_t w = ::pwrite(fd, p, n, (off_t)off);
        if (w <= 0) return;  // disk full etc.: like the reference's unchecked ofstream, the file is just short
        p += w;
        n -= (size_t)w;
        off += (size_t)w;
    }
}

// one z-plane of one array as text (pass 0: rho, pass 1: v); `out` is scratch storage that only ever grows,
// the return value is the number of bytes written to its beginning
template <typename T>
size_t format_plane(std::string &out, size_t dim, int pass, size_t z, const T *rho, const T *u)
{
    const size_t from = 1, to = dim - 1, n = dim * dim * dim;
    // formatted straight into the string's storage: at most 25 characters per value + one newline per row
    const size_t w = to - from;
    const size_t bound = w * (w * (pass == 0 ? 1 : 3) * 25 + 1) + 32;
    if (out.size() < bound) out.resize(bound);
    char *const begin = &out[0];
    char *p = begin;
    for (size_t y = from; y < to; ++y) {
        const size_t row = y * dim + z * dim * dim;
        if (pass == 0) {
            for (size_t x = from; x < to; ++x) {
                p += lbm_fmt::fmt_e16((double)rho[row + x], p);
                *p++ = ' ';
            }
        } else {
            for (size_t x = from; x < to; ++x) {
                for (size_t c = 0; c < 3; ++c) {
                    p += lbm_fmt::fmt_e16((double)u[c * n + row + x], p);
                    *p++ = ' ';
                }
            }
        }
        *p++ = '\n';
    }
    return (size_t)(p - begin);
}

// One array of the file, starting at byte `base`; returns the new end of the file.
template <typename T>
size_t write_array(const Sink &fd, size_t base, size_t dim, int pass, const T *rho, const T *u, unsigned nthreads)
{
    const size_t planes = dim - 2;
    const size_t rounds = (planes + nthreads - 1) / nthreads;
    std::vector<std::string> text(nthreads);
    std::vector<size_t> size(nthreads, 0);
    Barrier bar(nthreads);
    size_t end = base;  // every worker advances its own copy of the base identically; worker 0 reports it
    auto work = [&](unsigned j) {
        size_t my_base = base;
        for (size_t r = 0; r < rounds; ++r) {
            const size_t k = r * nthreads + j;
            size[j] = k < planes ? format_plane<T>(text[j], dim, pass, 1 + k, rho, u) : 0;
            bar.wait();  // every size of this round is known
            size_t before = 0, total = 0;
            for (unsigned i = 0; i < nthreads; ++i) {
                if (i < j) before += size[i];
                total += size[i];
            }
            write_at(fd, text[j].data(), size[j], my_base + before);
            my_base += total;
            bar.wait();  // everybody has read the sizes: they may be overwritten
        }
        if (j == 0) end = my_base;
    };
    std::vector<std::thread> pool;
    for (unsigned j = 1; j < nthreads; ++j) pool.emplace_back(work, j);
    work(0);
    for (auto &t : pool) t.join();
    return end;
}

// rho[N], u[3][N] in the reference's global layouts (kernels.cl:67-70).  nthreads == 0: one per core, at most
// 32.  A path that cannot be opened is a silent no-op, as in the reference (no check on its ofstream).
template <typename T>
void write_vti(const std::string &path, size_t dim, const T *rho, const T *u, unsigned nthreads)
{
    Sink fd;
    fd.fd = ::open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (fd.fd < 0) return;
    const size_t extent = dim - 3;
    const char *type = std::is_same<T, float>::value ? "Float32" : "Float64";
    char head[1024];
    int n = std::snprintf(head, sizeof head,
                          "<?xml version=\"1.0\"?>\n"
                          "<VTKFile type=\"ImageData\" version=\"0.1\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n"
                          "  <ImageData WholeExtent=\"0 %zu 0 %zu 0 %zu\" Origin=\"0 0 0\" Spacing=\"1 1 1\">\n"
                          "    <Piece Extent=\"0 %zu 0 %zu 0 %zu\">\n"
                          "      <PointData Scalars=\"rho\">\n"
                          "        <DataArray type=\"%s\" Name=\"rho\" NumberOfComponents=\"1\" format=\"ascii\">\n",
                          extent, extent, extent, extent, extent, extent, type);
    // upper bound of the file size: at most 25 characters per value ("-d.dddddddddddddddde-ddd "), one
    // newline per row, the XML around it
    {
        const size_t w = dim - 2;
        const size_t bound = 4096 + w * w * (w * 4 * 25 + 2);
        // a store into a mapping of a full file system raises SIGBUS where write() just comes back short (the
        // reference's unchecked ofstream would leave a short file): map only with room to spare
        struct statvfs vfs;
        const bool roomy = std::getenv("LBM_VTI_NO_MMAP") == nullptr && ::fstatvfs(fd.fd, &vfs) == 0 &&
                           (double)vfs.f_bavail * (double)vfs.f_frsize > 2.0 * (double)bound + (double)(64u << 20);
        if (roomy && ::ftruncate(fd.fd, (off_t)bound) == 0) {
            void *m = ::mmap(nullptr, bound, PROT_READ | PROT_WRITE, MAP_SHARED, fd.fd, 0);
            if (m != MAP_FAILED) {
                fd.map = static_cast<char *>(m);
                fd.map_size = bound;
            } else if (::ftruncate(fd.fd, 0) != 0) {
                // cannot happen on a file we just created; the pwrite path below still works
            }
        }
    }
    size_t pos = 0;
    write_at(fd, head, (size_t)n, pos);
    pos += (size_t)n;
    if (nthreads == 0) {
        nthreads = std::thread::hardware_concurrency();
        if (nthreads == 0) nthreads = 1;
        if (nthreads > 32) nthreads = 32;
    }
    if (nthreads > dim - 2) nthreads = (unsigned)(dim - 2);
    pos = write_array<T>(fd, pos, dim, 0, rho, u, nthreads);
    n = std::snprintf(head, sizeof head,
                      "        </DataArray>\n"
                      "        <DataArray type=\"%s\" Name=\"v\" NumberOfComponents=\"3\" format=\"ascii\">\n",
                      type);
    write_at(fd, head, (size_t)n, pos);
    pos += (size_t)n;
    pos = write_array<T>(fd, pos, dim, 1, rho, u, nthreads);
    const char tail[] = "        </DataArray>\n      </PointData>\n    </Piece>\n  </ImageData>\n</VTKFile>\n";
    write_at(fd, tail, sizeof tail - 1, pos);
    pos += sizeof tail - 1;
    if (fd.map) {
        ::munmap(fd.map, fd.map_size);
        if (::ftruncate(fd.fd, (off_t)pos) != 0) {
            // the file keeps its zero padding; nothing sensible to do (the reference checks no I/O either)
        }
    }
    ::close(fd.fd);
}

}  // namespace lbm_vti
