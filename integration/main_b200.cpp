// main_b200.cpp — builds the QuartetScores command line (-r/-e/-o/-q/-s/-t/-v) on top of libqscuda WITHOUT
// modifying or copying the reference: the reference's own main (src/QuartetScores.cpp) is compiled as is, from
// where it lies, and only the class it instantiates is swapped.
//
//   1. our QuartetScoreComputerB200.hpp defines qsb200::QuartetScoreComputer<CINT> (C-ABI calls into libqscuda.so);
//   2. the reference header is read once with its class renamed, so its `#pragma once` makes the main's own
//      #include "QuartetScoreComputer.hpp" a no-op;
//   3. `using qsb200::QuartetScoreComputer` makes the main's QuartetScoreComputer<uint8_t/16/32/64> resolve to ours.
//
// Build recipe: oracle/Makefile target `ref` (needs /root/reference; output oracle/_ref/QuartetScoresB200).
#include "QuartetScoreComputerB200.hpp"

#define QuartetScoreComputer QuartetScoreComputer_reference_cpu_unused
#include "QuartetScoreComputer.hpp"      // the reference's header (from -I$(REF)/src), class renamed
#undef QuartetScoreComputer
using qsb200::QuartetScoreComputer;

#include "QuartetScores.cpp"             // the reference's main, unmodified (from -I$(REF)/src)
