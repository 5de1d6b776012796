// `seeksv` executable: same command line as the reference binary (seeksv.cpp:26), all work behind the C ABI.
#include "../../include/seeksv_b200.h"
int main(int argc, char **argv) { return svb_main(argc, argv); }
