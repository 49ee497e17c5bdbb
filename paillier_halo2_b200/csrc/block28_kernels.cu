// block28_kernels.cu — placeholder until the block28 engine lands (next commit).
#include "engine.hpp"
namespace pb200 {
struct Block28Key { int dummy; };
Block28Key* block28_create(const BigInt&, const BigInt&, uint32_t, int, cudaStream_t, std::string* why, cudaError_t* e) {
    if (why) *why = "block28 engine not built"; if (e) *e = cudaSuccess; return nullptr;
}
void block28_destroy(Block28Key*) {}
const char* block28_name(const Block28Key*) { return "block28"; }
cudaError_t block28_encrypt(Block28Key*, const u64*, const u64*, size_t, u64*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t block28_tally(Block28Key*, const u64*, size_t, u64*, cudaStream_t) { return cudaErrorNotSupported; }
}
