/* TEST INFRASTRUCTURE — stand-in for the Bison-generated header that
 * src/nip.c:34 includes.  The reader that provides yyparse() in the
 * oracle/_ref build is refbuild/hugin_rd_parser.c. */
#ifndef HUGINNET_TAB_STANDIN_H
#define HUGINNET_TAB_STANDIN_H
int yyparse(void);
#endif
