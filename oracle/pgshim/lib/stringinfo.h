/* pgshim/lib/stringinfo.h -- intentionally minimal (see pgshim/postgres.h). */
#include "postgres.h"
#ifndef NDB_PGSHIM_STRINGINFO_H
#define NDB_PGSHIM_STRINGINFO_H
typedef struct StringInfoData { char *data; int len; int maxlen; int cursor; } StringInfoData;
typedef StringInfoData *StringInfo;
extern void initStringInfo(StringInfo s);
extern void appendStringInfo(StringInfo s, const char *fmt, ...);
extern void appendStringInfoString(StringInfo s, const char *str);
extern void appendStringInfoChar(StringInfo s, char c);
extern void resetStringInfo(StringInfo s);
#endif
