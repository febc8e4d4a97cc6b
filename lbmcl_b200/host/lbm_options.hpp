// Command-line contract of the LBMCL host program, restated for the B200 build.
//
// Same options, defaults, help text and failure behaviour as the reference parser
// (reference libs/lbm_options.hpp:32-50 defaults, :52-71 help, :74-196 parsing):
//   -P platform (accepted, ignored: there is no OpenCL platform)   -D device = CUDA ordinal
//   -d dim  -n viscosity  -u velocity  -i iterations  -e every  -w x,y,z  -s stride
//   -F double  -o optimize  -v vtk_path  -p dump_path  -m dump_map  -f dump_f  -h help
// plus two new options: -G / --gpus N splits the cube into N z-slabs over N devices, and
// -A / --aa selects the in-place AA-pattern kernels (one lattice, half the device memory).
// Unknown options and -h print the help and exit with status 1; an invalid number prints the
// reference's message for that option and exits with status 1.
#pragma once

#include <getopt.h>

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#define LBM_RESULTS_FOLDER "./results"

struct lbm_options {
    int platformID = -1;
    int deviceID = -1;
    size_t dim = 8;
    double viscosity = 0.0089;
    double velocity = 0.05;
    size_t iterations = 10;
    size_t every = 1;
    std::string vtk_path = LBM_RESULTS_FOLDER;
    size_t lwx = 8, lwy = 8, lwz = 8;
    size_t stride = 32;
    bool use_double = false;
    bool optimize = false;
    std::string dump_path = LBM_RESULTS_FOLDER;
    bool dump_map = false;
    bool dump_f = false;
    int gpus = 1;
    bool aa = false;

    static void print_help()
    {
        static const char *const lines[] = {
            "-P  --platform            Use the specified platform                     ",
            "-D  --device              Use the specified device                       ",
            "-d  --dim                 Set the lattice cube dimension                 ",
            "-n  --viscosity           Set the fluid viscosity                        ",
            "-u  --velocity            Set the x velocity of the moving wall          ",
            "-i  --iterations          Specify the number of iterations               ",
            "-e  --every               Save simulation results every N iterations     ",
            "-w  --work_group_size     Specify the work group size \"x,y,z\"          ",
            "-s  --stride              Specify the stride used in CSoA memory layout  ",
            "-F  --use_double          Make use of \"double\" type                    ",
            "-o  --optimize            Use \"cl-fast-relaxed-math\" in OpenCL kernels ",
            "-v  --vtk_path            Specify where store VTI files                  ",
            "-p  --dump_path           Specify where store dumps                      ",
            "-m  --dump_map            Dump the lattice map                           ",
            "-f  --dump_f              Dump the lattice \"f\" for each iteration      ",
            "-G  --gpus                Split the cube into z-slabs over N GPUs        ",
            "-A  --aa                  Use the in-place AA-pattern kernels            ",
            "-h  --help                Show this help message and exit                ",
        };
        for (const char *l : lines) std::cout << l << "\n";
        std::exit(1);
    }

    // non-negative integer or the reference's complaint for that option
    static long parse_count(const char *text, const char *complaint)
    {
        char *end = nullptr;
        const long v = std::strtol(text, &end, 10);
        if (end == text || v < 0) {
            std::cerr << complaint << std::endl;
            std::exit(1);
        }
        return v;
    }

    void process_args(int argc, char *argv[])
    {
        static const option long_opts[] = {
            {"platform", required_argument, nullptr, 'P'},   {"device", required_argument, nullptr, 'D'},
            {"dim", required_argument, nullptr, 'd'},        {"viscosity", required_argument, nullptr, 'n'},
            {"velocity", required_argument, nullptr, 'u'},   {"iterations", required_argument, nullptr, 'i'},
            {"every", required_argument, nullptr, 'e'},      {"work_group_size", required_argument, nullptr, 'w'},
            {"stride", required_argument, nullptr, 's'},     {"use_double", no_argument, nullptr, 'F'},
            {"optimize", no_argument, nullptr, 'o'},         {"vtk_path", required_argument, nullptr, 'v'},
            {"dump_path", required_argument, nullptr, 'p'},  {"dump_map", no_argument, nullptr, 'm'},
            {"dump_f", no_argument, nullptr, 'f'},           {"gpus", required_argument, nullptr, 'G'},
            {"aa", no_argument, nullptr, 'A'},               {"help", no_argument, nullptr, 'h'},
            {nullptr, 0, nullptr, 0}};
        opterr = 0;
        int opt;
        while ((opt = getopt_long(argc, argv, "P:D:d:n:u:i:e:v:w:s:Fop:mfhG:A", long_opts, nullptr)) >= 0) {
            switch (opt) {
                case 'P': platformID = (int)parse_count(optarg, "Please enter a valid platform"); break;
                case 'D': deviceID = (int)parse_count(optarg, "Please enter a valid device"); break;
                case 'd': dim = (size_t)parse_count(optarg, "Please enter a valid lattice dimension"); break;
                case 'n': {
                    const double v = std::atof(optarg);
                    if (v < 0) {
                        std::cerr << "Please enter a valid viscosity value" << std::endl;
                        std::exit(1);
                    }
                    viscosity = v;
                    break;
                }
                case 'u': velocity = std::atof(optarg); break;
                case 'i': iterations = (size_t)parse_count(optarg, "Please enter a valid number of iterations"); break;
                case 'e':
                    every = (size_t)parse_count(
                        optarg, "Please enter a valid number for save simulation results every N iterations");
                    break;
                case 'w': std::sscanf(optarg, "%zu,%zu,%zu", &lwx, &lwy, &lwz); break;
                case 's': stride = (size_t)parse_count(optarg, "Please enter a valid number for stride value"); break;
                case 'F': use_double = true; break;
                case 'o': optimize = true; break;
                case 'v':
                    vtk_path = optarg;
                    if (vtk_path.empty()) {
                        vtk_path = LBM_RESULTS_FOLDER;
                        std::cout << "VTI files will be stored in:" << vtk_path << std::endl;
                    }
                    break;
                case 'p':
                    dump_path = optarg;
                    if (dump_path.empty()) {
                        dump_path = LBM_RESULTS_FOLDER;
                        std::cout << "dump files will be stored in: " << dump_path << std::endl;
                    }
                    break;
                case 'm': dump_map = true; break;
                case 'f': dump_f = true; break;
                case 'A': aa = true; break;
                case 'G': gpus = (int)parse_count(optarg, "Please enter a valid number of GPUs"); break;
                default: print_help(); break;  // 'h', '?'
            }
        }
        if (gpus < 1) gpus = 1;
    }
};
