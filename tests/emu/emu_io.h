// Named-array files between the Python tests and the emulation harnesses (TEST INFRASTRUCTURE).
// Layout: u32 n | n x { u32 name_len, name, u32 elem_size, u64 count, data }.  tests/emu_io.py is the Python side.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <map>
#include <string>
#include <vector>

struct emu_array {
  uint32_t elem_size = 0;
  std::vector<char> bytes;
  template <class T>
  T* as() {
    if (bytes.empty()) return nullptr;
    if (sizeof(T) != elem_size) {
      fprintf(stderr, "emu_io: element size mismatch\n");
      exit(2);
    }
    return reinterpret_cast<T*>(bytes.data());
  }
  size_t count() const { return elem_size ? bytes.size() / elem_size : 0; }
};
typedef std::map<std::string, emu_array> emu_arrays;

inline emu_arrays emu_read(const char* path) {
  emu_arrays out;
  FILE* f = fopen(path, "rb");
  if (!f) {
    fprintf(stderr, "emu_io: cannot open %s\n", path);
    exit(2);
  }
  uint32_t n = 0;
  if (fread(&n, 4, 1, f) != 1) exit(2);
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t len = 0, es = 0;
    uint64_t cnt = 0;
    if (fread(&len, 4, 1, f) != 1) exit(2);
    std::string name(len, ' ');
    if (len && fread(&name[0], 1, len, f) != len) exit(2);
    if (fread(&es, 4, 1, f) != 1 || fread(&cnt, 8, 1, f) != 1) exit(2);
    emu_array a;
    a.elem_size = es;
    a.bytes.resize((size_t)es * cnt);
    if (!a.bytes.empty() && fread(a.bytes.data(), 1, a.bytes.size(), f) != a.bytes.size()) exit(2);
    out[name] = std::move(a);
  }
  fclose(f);
  return out;
}

template <class T>
inline emu_array emu_make(size_t count, T fill) {
  emu_array a;
  a.elem_size = sizeof(T);
  a.bytes.resize(sizeof(T) * count);
  T* p = reinterpret_cast<T*>(a.bytes.data());
  for (size_t i = 0; i < count; ++i) p[i] = fill;
  return a;
}

inline void emu_write(const char* path, const emu_arrays& arrays) {
  FILE* f = fopen(path, "wb");
  if (!f) exit(2);
  uint32_t n = (uint32_t)arrays.size();
  fwrite(&n, 4, 1, f);
  for (const auto& kv : arrays) {
    uint32_t len = (uint32_t)kv.first.size(), es = kv.second.elem_size;
    uint64_t cnt = kv.second.count();
    fwrite(&len, 4, 1, f);
    fwrite(kv.first.data(), 1, len, f);
    fwrite(&es, 4, 1, f);
    fwrite(&cnt, 8, 1, f);
    if (cnt) fwrite(kv.second.bytes.data(), 1, kv.second.bytes.size(), f);
  }
  fclose(f);
}
