/*
 * glsl_dialect.hpp -- the dialect macros the reference's shader sources are
 * written against (what renderer/src/shaders/glsl.glsl, metal.glsl, hlsl.glsl
 * provide for real shader compilers), for the C++ environment of glsl_env.hpp.
 * TEST INFRASTRUCTURE; written for this repo.
 *
 * Resources are namespace-scope objects the harness binds before calling a
 * shader's main; varyings and pixel-local-storage planes are namespace-scope
 * variables (the vertex main writes varyings, the harness interpolates, the
 * fragment main reads them). Both VERTEX and FRAGMENT are defined, as the
 * reference's Metal build does.
 */
#pragma once

#define VERTEX
#define FRAGMENT

#define INLINE inline
#define OUT(ARG_TYPE) ARG_TYPE&
#define INOUT(ARG_TYPE) ARG_TYPE&
#define MUL(A, B) ((A) * (B))

#define UNIFORM_BLOCK_BEGIN(IDX, NAME)                                                                                \
    struct NAME                                                                                                       \
    {
#define UNIFORM_BLOCK_END(NAME)                                                                                       \
    }                                                                                                                 \
    NAME;

#define ATTR_BLOCK_BEGIN(NAME)                                                                                        \
    struct NAME                                                                                                       \
    {
#define ATTR(IDX, TYPE, NAME) TYPE NAME
#define ATTR_BLOCK_END                                                                                                \
    }                                                                                                                 \
    ;
#define ATTR_UNPACK(ID, attrs, NAME, TYPE) TYPE NAME = attrs.NAME

#define VARYING(IDX, TYPE, NAME) TYPE NAME
#define FLAT
#define NO_PERSPECTIVE
#define OPTIONALLY_FLAT
#define VARYING_BLOCK_BEGIN
#define VARYING_BLOCK_END
#define VARYING_INIT(NAME, TYPE)
#define VARYING_PACK(NAME)
#define VARYING_UNPACK(NAME, TYPE)

#define VERTEX_TEXTURE_BLOCK_BEGIN
#define VERTEX_TEXTURE_BLOCK_END
#define FRAG_TEXTURE_BLOCK_BEGIN
#define FRAG_TEXTURE_BLOCK_END
#define DYNAMIC_SAMPLER_BLOCK_BEGIN
#define DYNAMIC_SAMPLER_BLOCK_END
#define VERTEX_STORAGE_BUFFER_BLOCK_BEGIN
#define VERTEX_STORAGE_BUFFER_BLOCK_END
#define FRAG_STORAGE_BUFFER_BLOCK_BEGIN
#define FRAG_STORAGE_BUFFER_BLOCK_END

// Re-declarable (the vertex and fragment blocks of one pipeline may both name a texture);
// the harness defines the objects.
#define TEXTURE_RGBA32UI(SET, IDX, NAME) extern Texture2D<uint4> NAME
#define TEXTURE_RGBA8(SET, IDX, NAME) extern TextureRGBA8 NAME
#define TEXTURE_R16F(SET, IDX, NAME) extern TextureR16F NAME
#define TEXTURE_R16F_1D_ARRAY(SET, IDX, NAME) extern Texture1DArrayR16F NAME
#define SAMPLER_LINEAR(TEXTURE_IDX, NAME)
#define SAMPLER_DYNAMIC_IMAGE(NAME)
#define TEXTURE_SAMPLE_LOD(NAME, SAMPLER_NAME, COORD, LOD) NAME.sampleLod(COORD)
#define TEXTURE_SAMPLE_DYNAMIC_LOD(NAME, SAMPLER_NAME, COORD, LOD) NAME.sampleLod(COORD)
#define TEXTURE_SAMPLE_LOD_1D_ARRAY(NAME, SAMPLER_NAME, X, ARRAY_INDEX, ARRAY_INDEX_NORMALIZED, LOD)                  \
    NAME.sampleLod(X, ARRAY_INDEX)
#define TEXEL_FETCH(NAME, COORD) NAME.fetch(COORD)
#define TEXTURE_CONTEXT_DECL
#define TEXTURE_CONTEXT_FORWARD

#define STORAGE_BUFFER_U32x2(IDX, GLSL_STRUCT_NAME, NAME) extern Buffer<uint2> NAME
#define STORAGE_BUFFER_U32x4(IDX, GLSL_STRUCT_NAME, NAME) extern Buffer<uint4> NAME
#define STORAGE_BUFFER_F32x4(IDX, GLSL_STRUCT_NAME, NAME) extern Buffer<float4> NAME
#define STORAGE_BUFFER_LOAD4(NAME, I) NAME._values[I]
#define STORAGE_BUFFER_LOAD2(NAME, I) NAME._values[I]

#define VERTEX_CONTEXT_DECL
#define VERTEX_CONTEXT_UNPACK
#define FRAGMENT_CONTEXT_DECL
#define FRAGMENT_CONTEXT_UNPACK
#define CLIP_CONTEXT_FORWARD
#define CLIP_CONTEXT_UNPACK

// The sources open their own brace after *_MAIN(...) and close it after EMIT_*; the GLSL
// dialect's macros open one more scope, which EMIT_VERTEX closes. Same shape here.
#define VERTEX_MAIN(NAME, Attrs, attrs, _vertexID, _instanceID)                                                       \
    inline float4 NAME(const Attrs& attrs, int _vertexID, int _instanceID)                                            \
    {
#define EMIT_VERTEX(_pos)                                                                                             \
    return _pos;                                                                                                      \
    }
#define FRAG_DATA_MAIN(DATA_TYPE, NAME) inline DATA_TYPE NAME()
#define EMIT_FRAG_DATA(VALUE) return VALUE

// Pixel local storage: four planes per pixel, bound by the harness. The colour planes are
// RGBA8 images: a store quantises (unorm8, round to nearest), as imageStore to rgba8 does.
#define PLS_BLOCK_BEGIN
#define PLS_BLOCK_END
#define PLS_DECL4F(IDX, NAME) extern float4 NAME
#define PLS_DECLUI(IDX, NAME) extern uint NAME
#define PLS_LOAD4F(PLANE) PLANE
#define PLS_LOADUI(PLANE) PLANE
#define PLS_STORE4F(PLANE, VALUE) PLANE = through_unorm8(VALUE)
#define PLS_STOREUI(PLANE, VALUE) PLANE = (VALUE)
#define PLS_PRESERVE_4F(PLANE)
#define PLS_PRESERVE_UI(PLANE)
#define PLS_INTERLOCK_BEGIN
#define PLS_INTERLOCK_END
#define PLS_MAIN(NAME) inline void NAME()
#define EMIT_PLS
