/* embed_cubin.S — places the relocatable scaffold cubin (built from scaffold.cu) into .rodata so
 * the shared library is self-contained. SCAFFOLD_CUBIN is set by the Makefile. */
    .section .rodata
    .global vb200_scaffold_cubin
    .type vb200_scaffold_cubin, @object
    .balign 16
vb200_scaffold_cubin:
    .incbin SCAFFOLD_CUBIN
vb200_scaffold_cubin_end:
    .size vb200_scaffold_cubin, vb200_scaffold_cubin_end - vb200_scaffold_cubin
    .global vb200_scaffold_cubin_size
    .type vb200_scaffold_cubin_size, @object
    .balign 8
vb200_scaffold_cubin_size:
    .quad vb200_scaffold_cubin_end - vb200_scaffold_cubin
    .size vb200_scaffold_cubin_size, 8
    .section .note.GNU-stack,"",@progbits
