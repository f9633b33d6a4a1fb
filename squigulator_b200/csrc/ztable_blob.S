/* Embeds data/ztable_v3.bin (Z32[32768] ++ Z2[8192], float32 LE; scripts/make_ztable.py) into libsqg.so */
    .section .rodata
    .global sqg_ztable_blob
    .type   sqg_ztable_blob, @object
    .balign 64
sqg_ztable_blob:
    .incbin ZTABLE_PATH
    .size   sqg_ztable_blob, . - sqg_ztable_blob
    .section .note.GNU-stack,"",@progbits
