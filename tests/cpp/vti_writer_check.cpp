// lbm_vti::write_vti() against a plain restatement of the reference's writer (reference lbmcl.hpp:261-334:
// one ofstream, std::scientific << std::setprecision(16) per value) -- the files must be byte-identical for
// every thread count, float and double, NaN / negative / tiny / zero values.
// Usage: vti_writer_check <scratch dir>   (exit status 0 = identical)
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "../../lbmcl_b200/host/vti_writer.hpp"

template <typename T>
static void reference_writer(const std::string &path, size_t dim, const T *rho, const T *u)
{
    const size_t from = 1, to = dim - 1, extent = to - from - 1, n = dim * dim * dim;
    const std::string type = std::is_same<T, float>::value ? "Float32" : "Float64";
    std::ofstream vtk(path.c_str());
    vtk << "<?xml version=\"1.0\"?>\n"
        << "<VTKFile type=\"ImageData\" version=\"0.1\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n"
        << "  <ImageData WholeExtent=\"0 " << extent << " 0 " << extent << " 0 " << extent << "\" Origin=\"0 0 0\" Spacing=\"1 1 1\">\n"
        << "    <Piece Extent=\"0 " << extent << " 0 " << extent << " 0 " << extent << "\">\n"
        << "      <PointData Scalars=\"rho\">\n"
        << "        <DataArray type=\"" << type << "\" Name=\"rho\" NumberOfComponents=\"1\" format=\"ascii\">\n";
    for (size_t z = from; z < to; ++z)
        for (size_t y = from; y < to; ++y) {
            for (size_t x = from; x < to; ++x)
                vtk << std::scientific << std::setprecision(16) << rho[x + y * dim + z * dim * dim] << " ";
            vtk << "\n";
        }
    vtk << "        </DataArray>\n"
        << "        <DataArray type=\"" << type << "\" Name=\"v\" NumberOfComponents=\"3\" format=\"ascii\">\n";
    for (size_t z = from; z < to; ++z)
        for (size_t y = from; y < to; ++y) {
            for (size_t x = from; x < to; ++x) {
                const size_t id = x + y * dim + z * dim * dim;
                vtk << std::scientific << std::setprecision(16) << u[id] << " " << u[n + id] << " " << u[2 * n + id] << " ";
            }
            vtk << "\n";
        }
    vtk << "        </DataArray>\n      </PointData>\n    </Piece>\n  </ImageData>\n</VTKFile>\n";
}

static std::string slurp(const std::string &path)
{
    std::ifstream f(path.c_str(), std::ios::binary);
    std::stringstream s;
    s << f.rdbuf();
    return s.str();
}

template <typename T>
static int run(const std::string &dir, size_t dim, const char *tag)
{
    const size_t n = dim * dim * dim;
    std::vector<T> rho(n), u(3 * n);
    std::mt19937_64 rng(7 + dim);
    std::uniform_real_distribution<double> d(-1.0, 1.0);
    for (size_t i = 0; i < n; ++i) {
        rho[i] = (T)(1.0 + 0.05 * d(rng));
        for (int c = 0; c < 3; ++c) {
            const double r = d(rng);
            u[c * n + i] = (T)(std::fabs(r) < 0.2 ? 0.0 : (std::fabs(r) < 0.4 ? r * 1e-30 : r * 0.05));
        }
        if (i % 7 == 0) rho[i] = u[i] = u[n + i] = u[2 * n + i] = (T)NAN;  // non-fluid cells
    }
    const std::string ref = dir + "/ref_" + tag + ".vti";
    reference_writer<T>(ref, dim, rho.data(), u.data());
    const std::string want = slurp(ref);
    int bad = 0;
    for (unsigned threads : {1u, 2u, 3u, 7u, 32u, 0u}) {
        const std::string out = dir + "/got_" + tag + "_" + std::to_string(threads) + ".vti";
        lbm_vti::write_vti<T>(out, dim, rho.data(), u.data(), threads);
        if (slurp(out) != want) {
            std::cout << "MISMATCH " << tag << " dim " << dim << " threads " << threads << "\n";
            ++bad;
        }
    }
    return bad;
}

int main(int argc, char **argv)
{
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    int bad = 0;
    bad += run<float>(dir, 8, "f32_8");
    bad += run<double>(dir, 8, "f64_8");
    bad += run<float>(dir, 4, "f32_4");
    bad += run<float>(dir, 32, "f32_32");
    bad += run<double>(dir, 16, "f64_16");
    lbm_vti::write_vti<float>(dir + "/no/such/dir/x.vti", 4, nullptr, nullptr, 1);  // silent no-op
    std::cout << (bad == 0 ? "all files byte-identical" : "FAILED") << "\n";
    return bad == 0 ? 0 : 1;
}
