// Pre-order serialisation of a host scene graph (grammar: include/artb_tokens.h). Used to prove that
// host scene construction matches the reference's (the oracle glue emits the same stream from the
// reference's own graph) and as an inspection aid behind artb_scene_dump_graph().
#ifndef ARTB_HOST_GRAPH_DUMP_H
#define ARTB_HOST_GRAPH_DUMP_H

#include "scene.h"

#include <cstdint>
#include <vector>

namespace artb {

std::uint64_t fnv1a64(const void* data, size_t n, std::uint64_t h = 1469598103934665603ULL);
std::vector<double> dump_graph(const hittable_list& world);

} // namespace artb

#endif
