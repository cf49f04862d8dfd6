// TEST INFRASTRUCTURE ONLY (oracle/): a multiplexer over the reference's own sub-programs (declared at
// /root/reference/src/megagta.cpp:8-15): `buildlib` / `buildgraph` are the path under test, `denovo` / `findstart` / `search`
// are the stages that consume its files (downstream acceptance).
#include <stdio.h>
#include <string.h>
int build_lib(int argc, char **argv);
int build_graph(int argc, char **argv);
int main_assemble(int argc, char **argv);
int find_start(int argc, char **argv);
int search(int argc, char **argv);
int main(int argc, char **argv) {
    if (argc >= 2 && strcmp(argv[1], "buildlib") == 0) return build_lib(argc - 1, argv + 1);
    if (argc >= 2 && strcmp(argv[1], "buildgraph") == 0) return build_graph(argc - 1, argv + 1);
    if (argc >= 2 && strcmp(argv[1], "denovo") == 0) return main_assemble(argc - 1, argv + 1);
    if (argc >= 2 && strcmp(argv[1], "findstart") == 0) return find_start(argc - 1, argv + 1);
    if (argc >= 2 && strcmp(argv[1], "search") == 0) return search(argc - 1, argv + 1);
    fprintf(stderr, "usage: %s buildlib|buildgraph|denovo|findstart|search [options]\n", argv[0]);
    return 1;
}
