// Host-side walk of the reference's op decomposition: reproduces, stage by stage, the instruction
// stream `new OP(...)` generates in the reference (src/Operation.cpp) without materialising
// Instruction objects, and returns the per-stage / per-opcode counts ("the count contract").
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace hml {

struct StageCount {
  std::string label;   // stage label as embedded in the reference's Instruction::Name
  std::string opcode;  // NTT | INTT | MULT | BCONV_STEP2 | AUTO
  uint64_t limb_ops;   // one limb-op = N/batchSize Instruction objects
};

struct TraceShape {
  uint32_t N, batch_size, max_level, L, alpha;
  uint32_t bconv_high, bconv_width;  // Driver replication factor of BCONV groups (Driver.h:307-320)
  uint32_t batch_count() const { return N / batch_size; }  // reference src/InsGen.cpp:12
  uint32_t beta() const { return (L + alpha - 1) / alpha; }  // reference src/Operation.cpp:22
  uint32_t digit_size(uint32_t j) const {                    // reference src/Operation.cpp:108-114
    uint32_t rem = L - j * alpha;
    return rem > alpha ? alpha : rem;
  }
};

// Appends the KeySwitch constructor's stages in creation order (reference src/Operation.cpp:35-53).
void trace_keyswitch(const TraceShape &s, std::vector<StageCount> &out);
void trace_rescale(const TraceShape &s, std::vector<StageCount> &out);
// op in {hmult, hrotate, hadd, pmult, padd}; returns false for any other name
// (reference bench_micro24.cpp:49-51 prints an error and exits 0).
bool trace_op(const std::string &op, const TraceShape &s, std::vector<StageCount> &out, std::string &err);

}  // namespace hml
