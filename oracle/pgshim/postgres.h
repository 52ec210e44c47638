/*
 * pgshim/postgres.h -- a minimal stand-in for PostgreSQL's postgres.h, written for this
 * repo (NOT taken from PostgreSQL or from the reference).  It declares only what
 * NeuronDB/src/vector/vector_distance.c and vector_distance_simd.c need to compile
 * unmodified outside a server, so that the reference's literal operator arithmetic can
 * be called from tests as oracle/_ref/libndb_ref_distance*.so.
 *
 * ereport(ERROR, ...) longjmps to a handler installed by shim_support.c
 * (ndb_ref_call_*), which reports the failure as a return code.
 */
#ifndef NDB_PGSHIM_POSTGRES_H
#define NDB_PGSHIM_POSTGRES_H

#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <setjmp.h>

#define PG_VERSION_NUM 170000
#define FLEXIBLE_ARRAY_MEMBER

typedef int8_t int8;
typedef int16_t int16;
typedef int32_t int32;
typedef int64_t int64;
typedef uint8_t uint8;
typedef uint16_t uint16;
typedef uint32_t uint32;
typedef uint64_t uint64;
typedef float float4;
typedef double float8;
typedef size_t Size;
typedef uintptr_t Datum;
typedef unsigned int Oid;
typedef char *Pointer;
typedef struct varlena { char vl_len_[4]; char vl_dat[FLEXIBLE_ARRAY_MEMBER]; } bytea;
typedef struct varlena text;

typedef struct MemoryContextData *MemoryContext;
extern MemoryContext CurrentMemoryContext;
typedef struct RelationData *Relation;
typedef struct ItemPointerData { uint16 bi_hi, bi_lo, ip_posid; } ItemPointerData;
typedef ItemPointerData *ItemPointer;

/* elog / ereport */
#define DEBUG5 10
#define DEBUG4 11
#define DEBUG3 12
#define DEBUG2 13
#define DEBUG1 14
#define LOG 15
#define INFO 17
#define NOTICE 18
#define WARNING 19
#define ERROR 21
#define FATAL 22
#define PANIC 23

extern void ndb_shim_raise(int elevel, const char *msg);
extern const char *ndb_shim_fmt(const char *fmt, ...);
#define errcode(x) 0
#define errmsg(...) ndb_shim_fmt(__VA_ARGS__)
#define errdetail(...) ndb_shim_fmt(__VA_ARGS__)
#define errhint(...) ndb_shim_fmt(__VA_ARGS__)
#define ereport(elevel, rest) \
	do { const char *ndb__m = (const char *) (uintptr_t) ((0) , rest); ndb_shim_raise(elevel, ndb__m); } while (0)
#define elog(elevel, ...) ndb_shim_raise(elevel, ndb_shim_fmt(__VA_ARGS__))

#define ERRCODE_NULL_VALUE_NOT_ALLOWED 1
#define ERRCODE_DATA_EXCEPTION 2
#define ERRCODE_INVALID_PARAMETER_VALUE 3
#define ERRCODE_NUMERIC_VALUE_OUT_OF_RANGE 4
#define ERRCODE_INTERNAL_ERROR 5
#define ERRCODE_OUT_OF_MEMORY 6
#define ERRCODE_PROGRAM_LIMIT_EXCEEDED 7
#define ERRCODE_FEATURE_NOT_SUPPORTED 8
#define ERRCODE_INSUFFICIENT_RESOURCES 9
#define ERRCODE_DATA_CORRUPTED 10
#define ERRCODE_ARRAY_SUBSCRIPT_ERROR 11
#define ERRCODE_INVALID_TEXT_REPRESENTATION 12
#define ERRCODE_UNDEFINED_OBJECT 13
#define ERRCODE_DIVISION_BY_ZERO 14
#define ERRCODE_EXTERNAL_ROUTINE_EXCEPTION 15
#define ERRCODE_CONNECTION_FAILURE 16
#define ERRCODE_SYNTAX_ERROR 17
#define ERRCODE_INVALID_NAME 18
#define ERRCODE_IO_ERROR 19
#define ERRCODE_UNDEFINED_TABLE 20
#define ERRCODE_UNDEFINED_COLUMN 21
#define ERRCODE_DATATYPE_MISMATCH 22
#define ERRCODE_UNDEFINED_FUNCTION 23
#define ERRCODE_INSUFFICIENT_PRIVILEGE 24
#define ERRCODE_OBJECT_NOT_IN_PREREQUISITE_STATE 25
#define ERRCODE_CONFIGURATION_LIMIT_EXCEEDED 26
#define ERRCODE_QUERY_CANCELED 27
#define ERRCODE_DUPLICATE_OBJECT 28
#define ERRCODE_NO_DATA_FOUND 29
#define ERRCODE_INVALID_BINARY_REPRESENTATION 30
#define ERRCODE_STRING_DATA_RIGHT_TRUNCATION 31
#define ERRCODE_NUMERIC_VALUE_OUT_OF_RANGE_ 32

/* memory */
extern void *palloc(Size n);
extern void *palloc0(Size n);
extern void *repalloc(void *p, Size n);
extern void pfree(void *p);
extern void *MemoryContextAlloc(MemoryContext c, Size n);
extern void *MemoryContextAllocZero(MemoryContext c, Size n);
extern MemoryContext MemoryContextSwitchTo(MemoryContext c);
extern char *pstrdup(const char *s);

#define Min(a, b) ((a) < (b) ? (a) : (b))
#define Max(a, b) ((a) > (b) ? (a) : (b))
#define MAXALIGN(x) (((uintptr_t) (x) + 7) & ~((uintptr_t) 7))
#define Assert(x) ((void) 0)
#define PGDLLEXPORT
#define pg_attribute_unused() __attribute__((unused))
#define lengthof(a) (sizeof(a) / sizeof((a)[0]))
#define likely(x) __builtin_expect((x) != 0, 1)
#define unlikely(x) __builtin_expect((x) != 0, 0)
#define CHECK_FOR_INTERRUPTS() ((void) 0)

#define VARHDRSZ 4
#define SET_VARSIZE(p, len) (*(int32 *) (p) = (int32) (len))
#define VARSIZE(p) (*(int32 *) (p))
#define VARSIZE_ANY(p) VARSIZE(p)
#define VARSIZE_ANY_EXHDR(p) (VARSIZE(p) - VARHDRSZ)
#define VARDATA(p) ((char *) (p) + VARHDRSZ)
#define VARDATA_ANY(p) VARDATA(p)

#define PointerGetDatum(x) ((Datum) (uintptr_t) (x))
#define DatumGetPointer(x) ((Pointer) (uintptr_t) (x))
#define Int32GetDatum(x) ((Datum) (int32) (x))
#define DatumGetInt32(x) ((int32) (x))
#define BoolGetDatum(x) ((Datum) ((x) ? 1 : 0))
#define DatumGetBool(x) ((bool) ((x) != 0))
static inline Datum Float4GetDatum(float4 x) { union { float4 f; uint32 u; } v; v.f = x; return (Datum) v.u; }
static inline float4 DatumGetFloat4(Datum d) { union { float4 f; uint32 u; } v; v.u = (uint32) d; return v.f; }
static inline Datum Float8GetDatum(float8 x) { union { float8 f; uint64 u; } v; v.f = x; return (Datum) v.u; }
static inline float8 DatumGetFloat8(Datum d) { union { float8 f; uint64 u; } v; v.u = (uint64) d; return v.f; }

#endif
