/* pgshim/fmgr.h -- minimal fmgr stand-in (see pgshim/postgres.h). */
#ifndef NDB_PGSHIM_FMGR_H
#define NDB_PGSHIM_FMGR_H
#include "postgres.h"
#define FUNC_MAX_ARGS 100
typedef struct NullableDatum { Datum value; bool isnull; } NullableDatum;
typedef struct FunctionCallInfoBaseData
{
	void	   *flinfo;
	void	   *context;
	void	   *resultinfo;
	Oid			fncollation;
	bool		isnull;
	short		nargs;
	NullableDatum args[8];
} FunctionCallInfoBaseData;
typedef FunctionCallInfoBaseData *FunctionCallInfo;
typedef Datum (*PGFunction) (FunctionCallInfo fcinfo);
#define PG_FUNCTION_ARGS FunctionCallInfo fcinfo
#define PG_FUNCTION_INFO_V1(f) extern Datum f(PG_FUNCTION_ARGS)
#define PG_MODULE_MAGIC extern int ndb_shim_module_magic
#define PG_NARGS() (fcinfo->nargs)
#define PG_ARGISNULL(n) (fcinfo->args[n].isnull)
#define PG_GETARG_DATUM(n) (fcinfo->args[n].value)
#define PG_GETARG_POINTER(n) DatumGetPointer(PG_GETARG_DATUM(n))
#define PG_GETARG_INT32(n) DatumGetInt32(PG_GETARG_DATUM(n))
#define PG_GETARG_INT16(n) ((int16) PG_GETARG_DATUM(n))
#define PG_GETARG_BOOL(n) DatumGetBool(PG_GETARG_DATUM(n))
#define PG_GETARG_FLOAT4(n) DatumGetFloat4(PG_GETARG_DATUM(n))
#define PG_GETARG_FLOAT8(n) DatumGetFloat8(PG_GETARG_DATUM(n))
#define PG_GETARG_TEXT_PP(n) ((text *) PG_GETARG_POINTER(n))
#define PG_DETOAST_DATUM(d) ((struct varlena *) DatumGetPointer(d))
#define PG_DETOAST_DATUM_COPY(d) PG_DETOAST_DATUM(d)
#define PG_RETURN_DATUM(x) return (x)
#define PG_RETURN_POINTER(x) return PointerGetDatum(x)
#define PG_RETURN_INT32(x) return Int32GetDatum(x)
#define PG_RETURN_BOOL(x) return BoolGetDatum(x)
#define PG_RETURN_FLOAT4(x) return Float4GetDatum(x)
#define PG_RETURN_FLOAT8(x) return Float8GetDatum(x)
#define PG_RETURN_NULL() do { fcinfo->isnull = true; return (Datum) 0; } while (0)
#define PG_RETURN_VOID() return (Datum) 0
#endif
