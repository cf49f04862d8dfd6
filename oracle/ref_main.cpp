// TEST INFRASTRUCTURE ONLY (oracle/): a two-entry multiplexer over the reference's own
// `build_lib` / `build_graph` (declared at /root/reference/src/megagta.cpp:8-9), so the oracle
// binary needs only the translation units on the buildgraph path.
#include <stdio.h>
#include <string.h>
int build_lib(int argc, char **argv);
int build_graph(int argc, char **argv);
int main(int argc, char **argv) {
    if (argc >= 2 && strcmp(argv[1], "buildlib") == 0) return build_lib(argc - 1, argv + 1);
    if (argc >= 2 && strcmp(argv[1], "buildgraph") == 0) return build_graph(argc - 1, argv + 1);
    fprintf(stderr, "usage: %s buildlib|buildgraph [options]\n", argv[0]);
    return 1;
}
