// lbmcl: the LBMCL host program on the B200 path.  Flow and output follow the reference's main.cpp:7-51
// (parse -> set up -> print the configuration -> run -> wait -> timing report on stdout -> the
// ';'-separated statistics line on stderr); LBMCL<T> is implemented over liblbm_b200.so.
#include <iostream>

#include "cli_options.hpp"
#include "lbmb200.hpp"

namespace {

template <typename Real>
int simulate(const lbm_options &opt)
{
    LBMCL<Real> lbm(opt.dim, static_cast<Real>(opt.viscosity), static_cast<Real>(opt.velocity), opt.iterations,
                    opt.every, opt.vtk_path, opt.lwx, opt.lwy, opt.lwz, opt.stride, opt.optimize, opt.dump_path,
                    opt.dump_map, opt.dump_f, opt.gpus, opt.aa);
    lbm.setupSimulation(opt.platformID, opt.deviceID);
    lbm.printConfiguration();
    lbm.performSimulationAndWait();

    // the four report lines of main.cpp:30-33, right-aligned labels
    const struct {
        const char *label;
        double value;
        const char *unit;
    } report[] = {
        {"   Total time: ", lbm.totalTimeMS(), " ms"},
        {" Kernels time: ", lbm.kernelsTimeMS(), " ms"},
        {"  Total MLUPS: ", lbm.MLUPS(), " MLUPS"},
        {"Kernels MLUPS: ", lbm.kernelsMLUPS(), " MLUPS"},
    };
    for (const auto &r : report) std::cout << r.label << r.value << r.unit << std::endl;

    std::cerr << lbm.statistics(';');  // consumed by benchmark.sh
    return 0;
}

}  // namespace

int main(int argc, char **argv)
{
    lbm_options opt;
    opt.process_args(argc, argv);
    return opt.use_double ? simulate<double>(opt) : simulate<float>(opt);
}
