// gguf.hpp -- minimal GGUF v2/v3 reader (no ggml).  Container layout per the reference's reader/writer:
// /root/reference/ggml/src/ggml.c:19594-19697 (structs), 19751-20063 (gguf_init_from_file), 20541-20640 (writer).
// Little-endian only.  The file is mmap'ed; tensors are views into the mapping.  Never throws across the C ABI:
// parse() returns false and fills `err`.
#pragma once
#include <fcntl.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <map>
#include <string>
#include <vector>

namespace cb {

enum GgufType : uint32_t {
    GT_U8 = 0, GT_I8, GT_U16, GT_I16, GT_U32, GT_I32, GT_F32, GT_BOOL, GT_STR, GT_ARR, GT_U64, GT_I64, GT_F64, GT_COUNT
};

struct GgufKV {
    std::string key;
    uint32_t type = 0;
    uint32_t arr_type = 0;       // for arrays
    uint64_t arr_n = 0;
    const uint8_t* raw = nullptr;   // start of the value payload in the mapping (after the type field)
    size_t raw_len = 0;             // bytes of the value payload
    uint64_t u = 0;                 // scalar integer / bool value
    double f = 0;                   // scalar float value
    std::string s;                  // string value
    std::vector<std::string> strs;  // string array
};

struct GgufTensor {
    std::string name;
    uint32_t n_dims = 0;
    uint64_t ne[4] = {1, 1, 1, 1};  // ggml order: ne[0] fastest
    uint32_t type = 0;              // ggml type id
    uint64_t offset = 0;            // relative to data section
    const uint8_t* data = nullptr;
    size_t nbytes = 0;
    uint64_t nelements() const { return ne[0] * ne[1] * ne[2] * ne[3]; }
};

inline size_t ggml_type_block_bytes(uint32_t t) {
    switch (t) { case 0: return 4; case 1: return 2; case 2: return 18; case 3: return 20; case 6: return 22; case 7: return 24; case 8: return 34; default: return 0; }
}
inline size_t ggml_type_block_elems(uint32_t t) { return (t == 0 || t == 1) ? 1 : 32; }

class GgufFile {
public:
    ~GgufFile() { close_(); }
    bool parse(const char* path, std::string& err) {
        fd_ = ::open(path, O_RDONLY);
        if (fd_ < 0) { err = std::string("cannot open ") + path; return false; }
        struct stat st;
        if (fstat(fd_, &st) != 0 || st.st_size < 24) { err = "file too small to be GGUF"; return false; }
        size_ = (size_t)st.st_size;
        void* m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (m == MAP_FAILED) { err = "mmap failed"; return false; }
        base_ = (const uint8_t*)m;
        size_t off = 0;
        uint32_t magic = 0;
        if (!rd(off, magic) || magic != 0x46554747u) { err = "not a GGUF file (bad magic)"; return false; }
        if (!rd(off, version) || version < 2) { err = "unsupported GGUF version"; return false; }
        uint64_t n_t = 0, n_kv = 0;
        if (!rd(off, n_t) || !rd(off, n_kv) || n_t > (1u << 24) || n_kv > (1u << 24)) { err = "corrupt GGUF header"; return false; }
        kvs.resize(n_kv);
        for (auto& kv : kvs) {
            if (!rd_str(off, kv.key) || !rd(off, kv.type)) { err = "corrupt GGUF kv"; return false; }
            const size_t v0 = off;
            if (!rd_val(off, kv)) { err = "corrupt GGUF kv value for " + kv.key; return false; }
            kv.raw = base_ + v0;
            kv.raw_len = off - v0;
            index_[kv.key] = &kv - kvs.data();
        }
        tensors.resize(n_t);
        for (auto& t : tensors) {
            if (!rd_str(off, t.name) || !rd(off, t.n_dims) || t.n_dims < 1 || t.n_dims > 4) { err = "corrupt GGUF tensor info"; return false; }
            uint64_t total = 1;
            for (uint32_t i = 0; i < t.n_dims; i++) {
                if (!rd(off, t.ne[i]) || t.ne[i] == 0 || t.ne[i] > (1ull << 40)) { err = "corrupt GGUF tensor info (dimension)"; return false; }
                if (total > (1ull << 48) / t.ne[i]) { err = "corrupt GGUF tensor info (element count overflows)"; return false; }
                total *= t.ne[i];
            }
            if (!rd(off, t.type) || !rd(off, t.offset)) { err = "corrupt GGUF tensor info"; return false; }
        }
        alignment = 32;
        if (const GgufKV* a = find("general.alignment")) alignment = (size_t)a->u;
        if (alignment == 0 || (alignment & (alignment - 1))) { err = "bad alignment"; return false; }
        meta_end = off;
        data_start = (off + alignment - 1) / alignment * alignment;
        for (size_t i = 0; i < tensors.size(); i++) {
            auto& t = tensors[i];
            const size_t bb = ggml_type_block_bytes(t.type), be = ggml_type_block_elems(t.type);
            if (!bb || t.ne[0] % be) { err = "tensor " + t.name + ": unsupported type or row length"; return false; }
            t.nbytes = (size_t)(t.nelements() / be) * bb;
            // overflow-safe bounds: offset and length are checked against what is LEFT of the file, never summed first
            if (data_start > size_ || t.offset > size_ - data_start || t.nbytes > size_ - data_start - t.offset) { err = "tensor " + t.name + " runs past end of file"; return false; }
            t.data = base_ + data_start + t.offset;
            tindex_[t.name] = i;
        }
        return true;
    }
    const GgufKV* find(const std::string& k) const { auto it = index_.find(k); return it == index_.end() ? nullptr : &kvs[it->second]; }
    const GgufTensor* tensor(const std::string& n) const { auto it = tindex_.find(n); return it == tindex_.end() ? nullptr : &tensors[it->second]; }

    uint32_t version = 0;
    size_t alignment = 32, data_start = 0, meta_end = 0;
    std::vector<GgufKV> kvs;
    std::vector<GgufTensor> tensors;

private:
    template <class T> bool rd(size_t& off, T& v) const {
        if (off > size_ || sizeof(T) > size_ - off) return false;
        memcpy(&v, base_ + off, sizeof(T)); off += sizeof(T); return true;
    }
    bool rd_str(size_t& off, std::string& s) const {
        uint64_t n = 0;
        if (!rd(off, n) || n > size_ - off) return false;
        s.assign((const char*)base_ + off, (size_t)n); off += (size_t)n; return true;
    }
    static size_t scalar_size(uint32_t t) {
        switch (t) { case GT_U8: case GT_I8: case GT_BOOL: return 1; case GT_U16: case GT_I16: return 2;
                     case GT_U32: case GT_I32: case GT_F32: return 4; case GT_U64: case GT_I64: case GT_F64: return 8; default: return 0; }
    }
    bool rd_scalar(size_t& off, uint32_t t, uint64_t& u, double& f) const {
        const size_t n = scalar_size(t);
        if (!n || off > size_ || n > size_ - off) return false;
        const uint8_t* p = base_ + off; off += n;
        switch (t) {
        case GT_U8: case GT_BOOL: u = *p; f = (double)u; break;
        case GT_I8: { int8_t v; memcpy(&v, p, 1); u = (uint64_t)(int64_t)v; f = v; } break;
        case GT_U16: { uint16_t v; memcpy(&v, p, 2); u = v; f = v; } break;
        case GT_I16: { int16_t v; memcpy(&v, p, 2); u = (uint64_t)(int64_t)v; f = v; } break;
        case GT_U32: { uint32_t v; memcpy(&v, p, 4); u = v; f = v; } break;
        case GT_I32: { int32_t v; memcpy(&v, p, 4); u = (uint64_t)(int64_t)v; f = v; } break;
        case GT_F32: { float v; memcpy(&v, p, 4); f = v; u = (uint64_t)v; } break;
        case GT_U64: { uint64_t v; memcpy(&v, p, 8); u = v; f = (double)v; } break;
        case GT_I64: { int64_t v; memcpy(&v, p, 8); u = (uint64_t)v; f = (double)v; } break;
        case GT_F64: { double v; memcpy(&v, p, 8); f = v; u = (uint64_t)v; } break;
        }
        return true;
    }
    bool rd_val(size_t& off, GgufKV& kv) const {
        if (kv.type == GT_STR) return rd_str(off, kv.s);
        if (kv.type == GT_ARR) {
            if (!rd(off, kv.arr_type) || !rd(off, kv.arr_n)) return false;
            if (kv.arr_type == GT_STR) {
                if (kv.arr_n > size_) return false;
                kv.strs.resize((size_t)kv.arr_n);
                for (auto& s : kv.strs) if (!rd_str(off, s)) return false;
                return true;
            }
            const size_t es = scalar_size(kv.arr_type);
            if (!es || kv.arr_n > (size_ - off) / es) return false;
            off += es * (size_t)kv.arr_n;
            return true;
        }
        return rd_scalar(off, kv.type, kv.u, kv.f);
    }
    void close_() {
        if (base_) munmap((void*)base_, size_);
        if (fd_ >= 0) ::close(fd_);
        base_ = nullptr; fd_ = -1;
    }
    int fd_ = -1;
    const uint8_t* base_ = nullptr;
    size_t size_ = 0;
    std::map<std::string, size_t> index_, tindex_;
};

}  // namespace cb
