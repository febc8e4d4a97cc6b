// Command-line contract of the LBMCL host program for the B200 build.
//
// Behaviour follows the reference parser (reference libs/lbm_options.hpp: defaults :32-50, help text
// :52-71, getopt loop :74-196): the same short and long options, the same defaults, the same complaint
// for an invalid number (exit status 1), help on -h or on an unknown option (exit status 1).
//   -P platform : accepted and ignored (there is no OpenCL platform on this path)
//   -D device   : CUDA ordinal
// Two options are new: -G/--gpus N (z-slab decomposition over N devices) and -A/--aa (in-place kernels).
// The option set lives in ONE table, from which the getopt strings and the help text are generated.
#pragma once

#include <getopt.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

struct lbm_options {
    int platformID = -1;
    int deviceID = -1;
    size_t dim = 8;
    double viscosity = 0.0089;
    double velocity = 0.05;
    size_t iterations = 10;
    size_t every = 1;
    std::string vtk_path = default_folder();
    size_t lwx = 8, lwy = 8, lwz = 8;
    size_t stride = 32;
    bool use_double = false;
    bool optimize = false;
    std::string dump_path = default_folder();
    bool dump_map = false;
    bool dump_f = false;
    int gpus = 1;
    bool aa = false;

    static const char *default_folder() { return "./results"; }

    struct spec {
        char key;
        const char *name;
        bool takes_value;
        const char *text;
    };
    static const std::vector<spec> &table()
    {
        static const std::vector<spec> t = {
            {'P', "platform", true, "Use the specified platform"},
            {'D', "device", true, "Use the specified device"},
            {'d', "dim", true, "Set the lattice cube dimension"},
            {'n', "viscosity", true, "Set the fluid viscosity"},
            {'u', "velocity", true, "Set the x velocity of the moving wall"},
            {'i', "iterations", true, "Specify the number of iterations"},
            {'e', "every", true, "Save simulation results every N iterations"},
            {'w', "work_group_size", true, "Specify the work group size \"x,y,z\""},
            {'s', "stride", true, "Specify the stride used in CSoA memory layout"},
            {'F', "use_double", false, "Make use of \"double\" type"},
            {'o', "optimize", false, "Use \"cl-fast-relaxed-math\" in OpenCL kernels"},
            {'v', "vtk_path", true, "Specify where store VTI files"},
            {'p', "dump_path", true, "Specify where store dumps"},
            {'m', "dump_map", false, "Dump the lattice map"},
            {'f', "dump_f", false, "Dump the lattice \"f\" for each iteration"},
            {'G', "gpus", true, "Split the cube into z-slabs over N GPUs"},
            {'A', "aa", false, "Use the in-place AA-pattern kernels"},
            {'h', "help", false, "Show this help message and exit"},
        };
        return t;
    }

    // "-d  --dim                 Set the lattice cube dimension                 " : the reference's columns
    [[noreturn]] static void print_help()
    {
        for (const spec &s : table()) {
            char line[160];
            std::snprintf(line, sizeof line, "-%c  --%-20s%-47s", s.key, s.name, s.text);
            std::cout << line << "\n";
        }
        std::exit(1);
    }

    [[noreturn]] static void reject(const char *what)
    {
        std::cerr << "Please enter a valid " << what << std::endl;
        std::exit(1);
    }
    static long count(const char *text, const char *what)
    {
        char *end = nullptr;
        const long v = std::strtol(text, &end, 10);
        if (end == text || v < 0) reject(what);
        return v;
    }

    void process_args(int argc, char *argv[])
    {
        std::string shorts;
        std::vector<option> longs;
        for (const spec &s : table()) {
            shorts += s.key;
            if (s.takes_value) shorts += ':';
            longs.push_back(option{s.name, s.takes_value ? required_argument : no_argument, nullptr, s.key});
        }
        longs.push_back(option{nullptr, 0, nullptr, 0});

        opterr = 0;
        for (int k; (k = getopt_long(argc, argv, shorts.c_str(), longs.data(), nullptr)) >= 0;) {
            switch (k) {
                case 'P': platformID = (int)count(optarg, "platform"); break;
                case 'D': deviceID = (int)count(optarg, "device"); break;
                case 'd': dim = (size_t)count(optarg, "lattice dimension"); break;
                case 'i': iterations = (size_t)count(optarg, "number of iterations"); break;
                case 'e': every = (size_t)count(optarg, "number for save simulation results every N iterations"); break;
                case 's': stride = (size_t)count(optarg, "number for stride value"); break;
                case 'G': gpus = (int)count(optarg, "number of GPUs"); break;
                case 'n':
                    viscosity = std::atof(optarg);
                    if (viscosity < 0) reject("viscosity value");
                    break;
                case 'u': velocity = std::atof(optarg); break;
                case 'w': std::sscanf(optarg, "%zu,%zu,%zu", &lwx, &lwy, &lwz); break;
                case 'v': set_folder(vtk_path, optarg, "VTI files will be stored in:"); break;
                case 'p': set_folder(dump_path, optarg, "dump files will be stored in: "); break;
                case 'F': use_double = true; break;
                case 'o': optimize = true; break;
                case 'm': dump_map = true; break;
                case 'f': dump_f = true; break;
                case 'A': aa = true; break;
                default: print_help();  // -h and anything unknown
            }
        }
        if (gpus < 1) gpus = 1;
    }

private:
    static void set_folder(std::string &dst, const char *value, const char *notice)
    {
        dst = value;
        if (dst.empty()) {
            dst = default_folder();
            std::cout << notice << dst << std::endl;
        }
    }
};
