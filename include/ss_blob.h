/* Header-only reader for the compiled-model blob (format: stretch_mujoco_b200/blob.py).
 * Pure container parsing -- no simulation arithmetic lives here.  Used by the C-ABI
 * library (csrc/) and by the CPU oracle (oracle/), which otherwise share no code. */
#ifndef SS_BLOB_H
#define SS_BLOB_H
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#pragma pack(push, 1)
typedef struct {
  char name[40];
  uint32_t dtype; /* 0=f64 1=i32 2=f32 3=u8 */
  uint32_t ndim;
  uint64_t shape[4];
  uint64_t offset;
  uint64_t nbytes;
} ss_blob_entry;
typedef struct {
  uint32_t objtype, count;
  uint64_t offset, nbytes;
} ss_blob_names;
#pragma pack(pop)

typedef struct {
  const unsigned char* base;
  size_t size;
  uint32_t narrays, ntables;
  const ss_blob_entry* entries;
  const ss_blob_names* tables;
} ss_blob;

static inline int ss_blob_open(ss_blob* b, const void* buf, size_t n) {
  if (n < 16 || memcmp(buf, "SSMBLOB1", 8) != 0) return -1;
  b->base = (const unsigned char*)buf;
  b->size = n;
  memcpy(&b->narrays, b->base + 8, 4);
  memcpy(&b->ntables, b->base + 12, 4);
  size_t need = 16 + (size_t)b->narrays * sizeof(ss_blob_entry) + (size_t)b->ntables * sizeof(ss_blob_names);
  if (need > n) return -1;
  b->entries = (const ss_blob_entry*)(b->base + 16);
  b->tables = (const ss_blob_names*)(b->base + 16 + (size_t)b->narrays * sizeof(ss_blob_entry));
  for (uint32_t i = 0; i < b->narrays; i++) {
    const ss_blob_entry* e = &b->entries[i];
    if (e->dtype > 3 || e->ndim > 4 || e->offset > n || e->nbytes > n - e->offset) return -1;   /* no offset + nbytes overflow */
  }
  for (uint32_t t = 0; t < b->ntables; t++) {
    const ss_blob_names* nt = &b->tables[t];
    if (nt->offset > n || nt->nbytes > n - nt->offset) return -1;
    /* every one of the `count` names must be NUL-terminated inside the table */
    uint32_t seen = 0;
    for (uint64_t k = 0; k < nt->nbytes && seen < nt->count; k++)
      if (b->base[nt->offset + k] == 0) seen++;
    if (seen < nt->count) return -1;
  }
  return 0;
}

static inline const ss_blob_entry* ss_blob_find(const ss_blob* b, const char* name) {
  for (uint32_t i = 0; i < b->narrays; i++)
    if (strncmp(b->entries[i].name, name, 40) == 0) return &b->entries[i];
  return NULL;
}

/* element count of an array, 0 if missing */
static inline size_t ss_blob_count(const ss_blob* b, const char* name) {
  const ss_blob_entry* e = ss_blob_find(b, name);
  if (!e) return 0;
  static const size_t isz[4] = {8, 4, 4, 1};
  return (size_t)(e->nbytes / isz[e->dtype & 3]);
}

static inline const double* ss_blob_f64(const ss_blob* b, const char* name) {
  const ss_blob_entry* e = ss_blob_find(b, name);
  return (e && e->dtype == 0) ? (const double*)(b->base + e->offset) : NULL;
}
static inline const int32_t* ss_blob_i32(const ss_blob* b, const char* name) {
  const ss_blob_entry* e = ss_blob_find(b, name);
  return (e && e->dtype == 1) ? (const int32_t*)(b->base + e->offset) : NULL;
}
static inline const float* ss_blob_f32(const ss_blob* b, const char* name) {
  const ss_blob_entry* e = ss_blob_find(b, name);
  return (e && e->dtype == 2) ? (const float*)(b->base + e->offset) : NULL;
}
static inline const unsigned char* ss_blob_u8(const ss_blob* b, const char* name) {
  const ss_blob_entry* e = ss_blob_find(b, name);
  return (e && e->dtype == 3) ? (const unsigned char*)(b->base + e->offset) : NULL;
}

/* name table lookup: returns pointer to the idx-th NUL-terminated name or NULL */
static inline const char* ss_blob_name(const ss_blob* b, uint32_t objtype, int idx) {
  for (uint32_t t = 0; t < b->ntables; t++) {
    if (b->tables[t].objtype != objtype) continue;
    if (idx < 0 || (uint32_t)idx >= b->tables[t].count) return NULL;
    const char* p = (const char*)(b->base + b->tables[t].offset);
    for (int i = 0; i < idx; i++) p += strlen(p) + 1;
    return p;
  }
  return NULL;
}
static inline int ss_blob_name2id(const ss_blob* b, uint32_t objtype, const char* name) {
  for (uint32_t t = 0; t < b->ntables; t++) {
    if (b->tables[t].objtype != objtype) continue;
    const char* p = (const char*)(b->base + b->tables[t].offset);
    for (uint32_t i = 0; i < b->tables[t].count; i++) {
      if (strcmp(p, name) == 0) return (int)i;
      p += strlen(p) + 1;
    }
  }
  return -1;
}
#endif
