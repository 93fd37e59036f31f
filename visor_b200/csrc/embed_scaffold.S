/* embed_scaffold.S — places the PTX of the kernel scaffolds (built from scaffold.cu) into .rodata so
 * the shared library is self-contained. SCAFFOLD_PTX is set by the Makefile. The runtime splices the
 * PTX of a pipeline's shader functions into this text and compiles the module with nvJitLink. */
    .section .rodata
    .global vb200_scaffold_ptx
    .type vb200_scaffold_ptx, @object
    .balign 16
vb200_scaffold_ptx:
    .incbin SCAFFOLD_PTX
vb200_scaffold_ptx_end:
    .byte 0
    .size vb200_scaffold_ptx, vb200_scaffold_ptx_end - vb200_scaffold_ptx
    .global vb200_scaffold_ptx_size
    .type vb200_scaffold_ptx_size, @object
    .balign 8
vb200_scaffold_ptx_size:
    .quad vb200_scaffold_ptx_end - vb200_scaffold_ptx
    .size vb200_scaffold_ptx_size, 8
    .section .note.GNU-stack,"",@progbits
