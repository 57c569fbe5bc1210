/* oracle-build stand-in: rna_data.cc:522 only needs TURN */
#ifndef TURN
#define TURN 3
#endif
