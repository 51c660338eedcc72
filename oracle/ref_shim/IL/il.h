/* oracle/ref_shim/IL/il.h -- TEST INFRASTRUCTURE.  Stands in for DevIL's <IL/il.h> (absent from the image) so that the
 * reference's src/Texture.cpp compiles where it lies: only the format / type enumerators that file switches on.  The values
 * are DevIL's (they mirror the OpenGL enumerators); nothing here loads an image. */
#ifndef REF_SHIM_IL_H
#define REF_SHIM_IL_H
typedef unsigned int ILenum;
typedef unsigned int ILuint;
typedef int ILint;
#define IL_BYTE             0x1400
#define IL_UNSIGNED_BYTE    0x1401
#define IL_SHORT            0x1402
#define IL_UNSIGNED_SHORT   0x1403
#define IL_INT              0x1404
#define IL_UNSIGNED_INT     0x1405
#define IL_FLOAT            0x1406
#define IL_DOUBLE           0x140A
#define IL_HALF             0x140B
#define IL_ALPHA            0x1906
#define IL_RGB              0x1907
#define IL_RGBA             0x1908
#define IL_LUMINANCE        0x1909
#define IL_LUMINANCE_ALPHA  0x190A
#define IL_BGR              0x80E0
#define IL_BGRA             0x80E1
#endif
