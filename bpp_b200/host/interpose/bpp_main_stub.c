/* bpp_main_stub.c -- entry point of the interposed bpp binaries (see locus_cuda.c): the reference's main() lives in
   libbppref.so under the name bpp_main (the build recipe compiles bpp.c with -Dmain=bpp_main). */
int bpp_main(int argc, char ** argv);
int main(int argc, char ** argv) { return bpp_main(argc, argv); }
