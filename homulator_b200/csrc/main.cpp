// Homulator.run — thin executable around hml_cli_main (see cli.cu).
#include "../../include/homulator_b200.h"

int main(int argc, char **argv) { return hml_cli_main(argc, argv); }
