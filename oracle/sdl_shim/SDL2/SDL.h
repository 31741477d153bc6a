/* Minimal stand-in for <SDL2/SDL.h> so the reference's src/camera.cpp compiles on a
 * headless box.  camera.cpp only polls the keyboard (src/camera.cpp:87-126); here no
 * key is ever down.  Like the real SDL.h (via SDL_stdinc.h) this pulls in <math.h>,
 * which under g++ makes the float overloads of sin/cos visible at global scope --
 * that decides which overload update_camera_vectors (src/camera.cpp:38-61) calls. */
#ifndef TRT_ORACLE_SDL_SHIM_H
#define TRT_ORACLE_SDL_SHIM_H
#include <math.h>
#include <stdint.h>
typedef uint8_t Uint8;
enum {
    SDL_SCANCODE_A = 4, SDL_SCANCODE_D = 7, SDL_SCANCODE_E = 8, SDL_SCANCODE_F = 9,
    SDL_SCANCODE_G = 10, SDL_SCANCODE_Q = 20, SDL_SCANCODE_R = 21, SDL_SCANCODE_S = 22,
    SDL_SCANCODE_T = 23, SDL_SCANCODE_W = 26, SDL_NUM_SCANCODES = 512
};
static inline const Uint8* SDL_GetKeyboardState(int* numkeys) {
    static Uint8 keys[SDL_NUM_SCANCODES];
    if (numkeys) *numkeys = SDL_NUM_SCANCODES;
    return keys;
}
#endif
