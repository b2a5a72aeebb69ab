// orb_matcher.cu -- ORBmatcher Hamming search on B200 (placeholder until the kernels land: every
// entry point fails loudly, there is no fallback).
#include "common.cuh"
using namespace orbb200;
struct orbb200_matcher { int dummy; };
#define NOT_YET(name) do { set_error(name ": not implemented yet"); return ORBB200_EINVAL; } while (0)
extern "C" int orbb200_matcher_create(int, int, int, orbb200_matcher**) { NOT_YET("orbb200_matcher_create"); }
extern "C" void orbb200_matcher_destroy(orbb200_matcher*) {}
extern "C" void* orbb200_matcher_stream(orbb200_matcher*) { return nullptr; }
extern "C" int orbb200_matcher_sync(orbb200_matcher*) { NOT_YET("orbb200_matcher_sync"); }
extern "C" int orbb200_matcher_last_launches(const orbb200_matcher*) { return 0; }
extern "C" int orbb200_descriptor_distance(orbb200_matcher*, const uint8_t*, const uint8_t*, int, int32_t*) { NOT_YET("orbb200_descriptor_distance"); }
extern "C" int orbb200_search_for_initialization(orbb200_matcher*, int, const orbb200_frame_view*, const orbb200_frame_view*, int, int, float, int, int, float*, int32_t*, int32_t*, int) { NOT_YET("orbb200_search_for_initialization"); }
extern "C" int orbb200_search_by_projection(orbb200_matcher*, int, const orbb200_frame_view*, const float*, const orbb200_mappoint_view*, int32_t*, const int32_t*, const float*, int, int, int, float, float, int32_t*, int) { NOT_YET("orbb200_search_by_projection"); }
