// mcpm_simulate <inputfile> [--q5-compat] [--quiet] [--out DIR]
// The reference's `simulate <inputfile>` (src/main.cpp:16-97) on the B200 engine.
#include <cstdio>
#include <cstring>

#include "../../../include/mcpm.h"

int main(int argc, char** argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <inputfile> [--q5-compat] [--quiet] [--out DIR]\n", argv[0]);
        return 2;
    }
    int flags = 0;
    const char* out = nullptr;
    for (int i = 2; i < argc; i++) {
        if (!std::strcmp(argv[i], "--q5-compat")) flags |= 1;
        else if (!std::strcmp(argv[i], "--quiet")) flags |= 2;
        else if (!std::strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
    }
    int rc = mcpm_simulate(argv[1], out, flags);
    if (rc != MCPM_OK) std::fprintf(stderr, "mcpm_simulate failed (%d): %s\n", rc, mcpm_last_error());
    return rc == MCPM_OK ? 0 : 1;
}
