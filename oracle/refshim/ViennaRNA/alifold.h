/* oracle-build stand-in (empty): alignment.cc includes it but uses nothing from it */
