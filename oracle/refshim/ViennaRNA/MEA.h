/* oracle-build stand-in: rna_data.cc:1773 calls MEA() for --consensus-structure mea only */
#include <ViennaRNA/data_structures.h>
#ifdef __cplusplus
extern "C"
#endif
float MEA(plist *p, char *structure, double gamma);
