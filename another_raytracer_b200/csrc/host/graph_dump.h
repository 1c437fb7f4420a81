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

// The inverse: rebuild a host graph from a token stream (a foreign host's graph serialised by
// include/artb_reference_adapter.hpp, or dump_graph's own output). The stream carries only the HASHES of Perlin tables
// and image texels; `blobs` supplies the bytes (noise: 256 x 3 doubles ranvec, then perm_x, perm_y, perm_z as 256 int32
// each = 9216 B; image: width*height*bytes_per_pixel). Equal material / texture token sequences become ONE shared object
// (the flattener deduplicates by pointer). Nothing here draws from the global RNG. Throws std::logic_error on a
// malformed stream or a missing / mis-sized blob.
struct GraphBlob { std::uint64_t hash; const void* data; std::uint64_t bytes; };
hittable_list build_graph(const double* tokens, size_t n_tokens, const GraphBlob* blobs, size_t n_blobs);

} // namespace artb

#endif
