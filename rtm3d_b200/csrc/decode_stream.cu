// placeholder until the streaming kernel lands
#include "params.h"
namespace rtm3d {
bool stream_eligible(const DecodeParams&, int, int) { return false; }
int launch_stream(const DecodeParams&, int, int, cudaStream_t) { return -1000; }
}
