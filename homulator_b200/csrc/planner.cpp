#include "planner.h"

#include <algorithm>

namespace hml {

void trace_keyswitch(const TraceShape &s, std::vector<StageCount> &out) {
  const uint64_t L = s.L, A = s.alpha, E = L + A, beta = s.beta();
  // K1 ModUpINTT: one INTT per input limb (reference src/Operation.cpp:63-102)
  out.push_back({"_ModUp_INTT(", "INTT", L});
  uint64_t decomp = 0, bconv_up = 0, ntt_up = 0;
  for (uint32_t j = 0; j < beta; ++j) {
    const uint64_t aj = s.digit_size(j);
    decomp += aj;                // K2 ModUpDecompFusionBConvStep1 (:104-135): a_j scalar multiplies
    bconv_up += aj * (E - aj);   // K3 ModUpBConvStep2 (:137-188): a_j MACs for each of the L+alpha-a_j outputs
    ntt_up += E;                 // K4 ModUpNTT (:190-292): ALL L+alpha limbs of the digit (delta D3)
  }
  out.push_back({"_decompFusionBConvStep1_beta(", "MULT", decomp});
  out.push_back({"_BCONVStep2_beta(", "BCONV_STEP2", bconv_up});
  out.push_back({"_Modup_NTT_beta(", "NTT", ntt_up});
  // K5 InnerProduceOperation (:294-414): per key component, beta==1 -> one pass, else beta-1 passes
  out.push_back({"_InnerProducOperation(", "MULT", 2 * E * std::max<uint64_t>(1, beta - 1)});
  // K6..K10 ModDown (:417-590).  K9 is emitted with opcode INTT (GenNTT(..., false, ...), :535-539) — delta D1.
  out.push_back({"ModDown_INTT(", "INTT", 2 * A});
  out.push_back({"_ModDownBConvStep1_Level(", "MULT", 2 * A});
  out.push_back({"_ModDownBConvStep2_", "BCONV_STEP2", 2 * A * L});
  out.push_back({"ModDown_NTT(", "INTT", 2 * L});
  out.push_back({"_KeySwitchFinalOutput_Level(", "MULT", 2 * L});
}

void trace_rescale(const TraceShape &s, std::vector<StageCount> &out) {
  // one polynomial (reference src/Operation.cpp:741-911): 1 INTT, 1 NTT (delta D2), L-1 sub, L-1 mul
  out.push_back({"_Rescale_INTT(", "INTT", 1});
  out.push_back({"_Rescale_NTT_level(", "NTT", 1});
  out.push_back({"_Rescale_Sub_Level(", "MULT", (uint64_t)s.L - 1});
  out.push_back({"_Rescale_Mul_Level(", "MULT", (uint64_t)s.L - 1});
}

bool trace_op(const std::string &op, const TraceShape &s, std::vector<StageCount> &out, std::string &err) {
  const uint64_t L = s.L;
  if (s.N == 0 || s.batch_size == 0 || s.N % s.batch_size != 0 || s.alpha == 0 || L == 0) {
    err = "invalid trace shape";
    return false;
  }
  if (op == "hmult") {
    if (L < 2) {
      err = "hmult needs currentLevel >= 2 (the reference's Rescale segfaults at 1)";
      return false;
    }
    // TensorCompute (reference src/Operation.cpp:592-739)
    out.push_back({"_TensorCompute_D0", "MULT", L});
    out.push_back({"_TensorCompute_D1", "MULT", L});
    out.push_back({"_TensorCompute_D2", "MULT", L});
    trace_keyswitch(s, out);
    out.push_back({"_HMULTHadd_Level(", "MULT", 2 * L});  // :967-1005
    std::vector<StageCount> r;
    trace_rescale(s, r);                                   // :1008-1022, once per output polynomial
    for (auto &st : r) { st.limb_ops *= 2; out.push_back(st); }
    return true;
  }
  if (op == "hrotate") {
    out.push_back({"_HADD_Level(", "AUTO", 2 * L});        // :1302-1319 (the label really says HADD)
    trace_keyswitch(s, out);
    out.push_back({"_HROTATE_HADD_Level(", "MULT", L});    // :1339-1357
    return true;
  }
  if (op == "hadd") { out.push_back({"_HADD_Level(", "MULT", 2 * L}); return true; }     // :1146-1170
  if (op == "pmult") { out.push_back({"_HMULT_level(", "MULT", 2 * L}); return true; }   // :1485-1515
  if (op == "padd") { out.push_back({"_PADD_Level(", "MULT", 2 * L}); return true; }     // :1650-1672
  err = "Error operation requirement, please double confirm!";
  return false;
}

}  // namespace hml
