// main.cpp — command line renderer with the reference's flags (-o -spp -j -bdpt,
// reference main.cpp:38-47) plus scene / resolution / model-directory selection.
// The scene scripts themselves are in tpt_host.cpp (tpth_scene_build); a user's own
// script is the reference main.cpp compiled against host/*.hpp.
#include <cstdio>
#include <sstream>
#include <string>

#include "tpt_host.h"

template <typename T> static T Arg(int argc, char** argv, const char* name, const T& dflt) {
    for (int i = 0; i + 1 < argc; i++)
        if (std::string(argv[i]) == name) {
            T v;
            std::stringstream ss(argv[i + 1]);
            ss >> v;
            return v;
        }
    return dflt;
}

int main(int argc, char** argv) {
    const std::string out = Arg(argc, argv, "-o", std::string("output.ppm"));
    const int spp = Arg(argc, argv, "-spp", 1);
    const int bdpt = Arg(argc, argv, "-bdpt", 1);
    const int full = Arg(argc, argv, "-ptfull", 0);
    const int w = Arg(argc, argv, "-w", 784), h = Arg(argc, argv, "-h", 784);
    const int device = Arg(argc, argv, "-device", 0);
    const std::string scene = Arg(argc, argv, "-scene", std::string("silver"));   // what main.cpp ships
    const std::string models = Arg(argc, argv, "-models", std::string("../models"));
    TpthScene* s = tpth_scene_build(scene.c_str(), models.c_str(), w, h);
    if (tpth_scene_error(s)) {
        std::fprintf(stderr, "%s\n", tpth_scene_error(s));
        return 1;
    }
    double sec = 0;
    const int gpus = Arg(argc, argv, "-gpus", 1);       // GPUs sharing the frame; -split 0 interleave 1 tile 2 spp 3 tile x spp
    const int split = Arg(argc, argv, "-split", 0);
    const int rc = gpus > 1 ? tpth_render_gpus(s, out.c_str(), spp, bdpt, full, gpus, split, nullptr, &sec, nullptr)
                            : tpth_render(s, out.c_str(), spp, bdpt, full, device, nullptr, &sec);
    if (rc == 0) std::printf("rendered %s %dx%d spp %d in %.3f s -> %s\n", scene.c_str(), w, h, spp, sec, out.c_str());
    tpth_scene_destroy(s);
    return rc;
}
