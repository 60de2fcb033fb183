// TEST INFRASTRUCTURE — measurement probe of the UNMODIFIED reference, not product code.
//
// Compiled together with the reference's own src/*.cpp where they lie under /root/reference
// (see oracle/Makefile, target _ref/count.run).  It constructs the reference's op object --
// all instruction generation happens in the constructor (reference src/Operation.cpp:913-1023
// for HMULT, :1271-1358 for HROTATE, :1114-1176 HADD, :1453-1523 PMULT, :1618-1680 PADD) --
// does NOT call simulate(), and tallies the Instruction objects the reference's Driver holds
// (reference include/Driver.h:14-16 sentInsFIFO, filled by dispatchInstructions :71-105) by
// opcode and by stage label.  The output is one JSON object on the last stdout line.
//
// Argument handling mirrors reference bench_test/bench_micro24.cpp:6-27.
#include <algorithm>
#include <cassert>
#include <cmath>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <chrono>
#include <stdio.h>
#include <typeinfo>

#define private public
#include "Basic.h"
#include "Operation.h"
#include "Arch.h"
#include "Driver.h"
#undef private

// Stage labels as they appear inside Instruction::Name (reference src/Operation.cpp, see
// SURVEY.md section 8a).  Order matters: first match wins.
static const char *kStageLabels[] = {
    "_TensorCompute_D0", "_TensorCompute_D1", "_TensorCompute_D2",
    "_ModUp_INTT(", "_decompFusionBConvStep1_beta(", "_BCONVStep2_beta(",
    "_Modup_NTT_beta(", "_InnerProducOperation(", "ModDown_INTT(",
    "_ModDownBConvStep1_Level(", "_ModDownBConvStep2_", "ModDown_NTT(",
    "_KeySwitchFinalOutput_Level(", "_HMULTHadd_Level(", "_Rescale_INTT(",
    "_Rescale_NTT_level(", "_Rescale_Sub_Level(", "_Rescale_Mul_Level(",
    "_HROTATE_HADD_Level(", "_HADD_Level(", "_HMULT_level(", "_PADD_Level(",
};

template <class OP>
static void tally(OP *op, Arch *arch, std::map<std::string, unsigned long long> &byOp,
                  std::map<std::string, unsigned long long> &byStage,
                  unsigned long long &driverTotal) {
  Driver *d = op->driver;
  for (uint32_t c = 0; c < d->cluster; c++) {
    for (auto &kv : d->sentInsFIFO[c]) {
      for (auto &group : kv.second) {
        for (Instruction *ins : group) {
          byOp[ins->GetOpName()]++;
          const std::string &nm = ins->Name;
          std::string st = "other";
          for (const char *lab : kStageLabels) {
            if (nm.find(lab) != std::string::npos) { st = lab; break; }
          }
          byStage[st + "|" + ins->GetOpName()]++;
        }
      }
    }
  }
  d->IssueInsFromDramToChip(arch);
  driverTotal = d->totalIns;
}

int main(int argc, char *argv[]) {
  if (argc < 6) {
    std::cerr << "Usage: " << argv[0] << " <cfg> <op> <maxLevel> <curLevel> <alpha> [cluster]\n";
    return 1;
  }
  std::string path = argv[1], ops = argv[2];
  // silence the reference's config dump / Malloc lines: keep only our JSON on stdout
  std::streambuf *old = std::cout.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf());
  auto t0 = std::chrono::steady_clock::now();
  Config *config = new Config(path);
  uint32_t maxlevel = std::atoi(argv[3]);
  uint32_t currentlevel = std::atoi(argv[4]);
  uint32_t alpha = std::atoi(argv[5]);
  if (argc > 6) config->setValue("cluster", std::atoi(argv[6]));
  Arch *arch = new Arch(config);
  std::map<std::string, unsigned long long> byOp, byStage;
  unsigned long long driverTotal = 0;
  if (ops == "hmult") {
    tally(new HMULT("test_hmult", maxlevel, currentlevel, alpha, config, arch), arch, byOp, byStage, driverTotal);
  } else if (ops == "hrotate") {
    tally(new HROTATE("test_hrotate", maxlevel, currentlevel, alpha, config, arch), arch, byOp, byStage, driverTotal);
  } else if (ops == "hadd") {
    tally(new HADD("test_hadd", maxlevel, currentlevel, alpha, config, arch), arch, byOp, byStage, driverTotal);
  } else if (ops == "pmult") {
    tally(new PMULT("test_pmult", maxlevel, currentlevel, alpha, config, arch), arch, byOp, byStage, driverTotal);
  } else if (ops == "padd") {
    tally(new PADD("test_ADD", maxlevel, currentlevel, alpha, config, arch), arch, byOp, byStage, driverTotal);
  } else {
    std::cout.rdbuf(old);
    std::cout << "{\"error\": \"unknown op\"}\n";
    return 0;
  }
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::cout.rdbuf(old);
  unsigned long long total = 0;
  for (auto &kv : byOp) total += kv.second;
  std::cout << "{\"op\": \"" << ops << "\", \"N\": " << config->getValue("N")
            << ", \"batchSize\": " << config->getValue("batchSize")
            << ", \"maxLevel\": " << maxlevel << ", \"L\": " << currentlevel
            << ", \"alpha\": " << alpha << ", \"total\": " << total
            << ", \"driverTotal\": " << driverTotal << ", \"insgen_seconds\": " << secs
            << ", \"by_opcode\": {";
  bool first = true;
  for (auto &kv : byOp) { std::cout << (first ? "" : ", ") << "\"" << kv.first << "\": " << kv.second; first = false; }
  std::cout << "}, \"by_stage\": {";
  first = true;
  for (auto &kv : byStage) { std::cout << (first ? "" : ", ") << "\"" << kv.first << "\": " << kv.second; first = false; }
  std::cout << "}}" << std::endl;
  return 0;
}
