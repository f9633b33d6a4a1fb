/* Embeds data/ztable_v2.bin (Z1[32768] ++ Z2[16384], float32 LE; scripts/make_ztable.py) into libsqg.so */
    .section .rodata
    .global sqg_ztable_blob
    .type   sqg_ztable_blob, @object
    .balign 64
sqg_ztable_blob:
    .incbin ZTABLE_PATH
    .size   sqg_ztable_blob, . - sqg_ztable_blob
    .section .note.GNU-stack,"",@progbits
