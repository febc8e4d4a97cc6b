// lbmcl: the LBMCL host program on the B200 path.  Same flow and output as the reference's
// main.cpp:7-51 -- parse, set up, print the configuration, run, wait, four timing lines on stdout,
// the ';'-separated statistics line on stderr -- with LBMCL<T> implemented over liblbm_b200.so.
#include <iostream>

#include "lbm_options.hpp"
#include "lbmb200.hpp"

template <typename T>
static void run(const lbm_options &o)
{
    LBMCL<T> sim(o.dim, (T)o.viscosity, (T)o.velocity, o.iterations, o.every, o.vtk_path, o.lwx, o.lwy, o.lwz,
                 o.stride, o.optimize, o.dump_path, o.dump_map, o.dump_f, o.gpus, o.aa);
    sim.setupSimulation(o.platformID, o.deviceID);
    sim.printConfiguration();
    sim.performSimulation();
    sim.waitCompletion();

    std::cout << "   Total time: " << sim.totalTimeMS() << " ms" << std::endl;
    std::cout << " Kernels time: " << sim.kernelsTimeMS() << " ms" << std::endl;
    std::cout << "  Total MLUPS: " << sim.MLUPS() << " MLUPS" << std::endl;
    std::cout << "Kernels MLUPS: " << sim.kernelsMLUPS() << " MLUPS" << std::endl;

    std::cerr << sim.statistics(';');
}

int main(int argc, char *argv[])
{
    lbm_options opts;
    opts.process_args(argc, argv);
    if (opts.use_double) run<double>(opts);
    else run<float>(opts);
    return 0;
}
