/*
 * Headless stand-in for the SDL-1.2 API surface that the fray reference uses.
 *
 * TEST INFRASTRUCTURE ONLY. This header lets the *unmodified* reference sources under
 * /root/reference/src compile and run without a display, so that the reference's own CPU renderer
 * can act as the parity oracle and the CPU baseline (see oracle/Makefile, DESIGN.md "Oracle").
 * Nothing in the product (fray_b200/, include/) includes or links this file.
 *
 * Only the ~40 symbols the reference touches are declared:
 * threads, mutexes and condition variables are thin pthread wrappers, the "screen" is a
 * malloc'd 32-bit surface, the event queue is always empty and SDL_WaitEvent reports SDL_QUIT
 * immediately so that `fray scene.fray` exits as soon as the frame is rendered.
 */
#ifndef FRAY_ORACLE_SDL_SHIM_H
#define FRAY_ORACLE_SDL_SHIM_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint8_t  Uint8;
typedef uint16_t Uint16;
typedef uint32_t Uint32;

#define SDL_INIT_VIDEO 0x20u
#define SDL_FULLSCREEN 0x80000000u

typedef struct SDL_PixelFormat {
	Uint8 BitsPerPixel, BytesPerPixel;
	Uint8 Rshift, Gshift, Bshift, Ashift;
} SDL_PixelFormat;

typedef struct SDL_Surface {
	Uint32 flags;
	SDL_PixelFormat* format;
	int w, h;
	Uint16 pitch;
	void* pixels;
} SDL_Surface;

/* keyboard / event subset */
typedef enum {
	SDLK_ESCAPE = 27,
	SDLK_KP2 = 258, SDLK_KP4 = 260, SDLK_KP6 = 262, SDLK_KP8 = 264,
	SDLK_UP = 273, SDLK_DOWN = 274, SDLK_RIGHT = 275, SDLK_LEFT = 276,
	SDLK_F12 = 293,
	SDLK_LAST = 323
} SDLKey;

typedef enum { KMOD_NONE = 0, KMOD_LSHIFT = 1, KMOD_RSHIFT = 2 } SDLMod;

enum { SDL_NOEVENT = 0, SDL_KEYDOWN = 2, SDL_MOUSEBUTTONDOWN = 5, SDL_QUIT = 12 };

typedef struct SDL_keysym { Uint8 scancode; SDLKey sym; SDLMod mod; Uint16 unicode; } SDL_keysym;
typedef struct SDL_KeyboardEvent { Uint8 type, which, state; SDL_keysym keysym; } SDL_KeyboardEvent;
typedef struct SDL_MouseButtonEvent { Uint8 type, which, button, state; Uint16 x, y; } SDL_MouseButtonEvent;
typedef union SDL_Event {
	Uint8 type;
	SDL_KeyboardEvent key;
	SDL_MouseButtonEvent button;
} SDL_Event;

int SDL_Init(Uint32 flags);
void SDL_Quit(void);
char* SDL_GetError(void);
SDL_Surface* SDL_SetVideoMode(int width, int height, int bpp, Uint32 flags);
int SDL_Flip(SDL_Surface* screen);
void SDL_UpdateRect(SDL_Surface* screen, int x, int y, Uint32 w, Uint32 h);
void SDL_WM_SetCaption(const char* title, const char* icon);
int SDL_ShowCursor(int toggle);
Uint32 SDL_GetTicks(void);
void SDL_Delay(Uint32 ms);
int SDL_PollEvent(SDL_Event* event);
int SDL_WaitEvent(SDL_Event* event);
Uint8* SDL_GetKeyState(int* numkeys);
Uint8 SDL_GetRelativeMouseState(int* x, int* y);

/* threads */
typedef struct SDL_Thread SDL_Thread;
SDL_Thread* SDL_CreateThread(int (*fn)(void*), void* data);
void SDL_WaitThread(SDL_Thread* thread, int* status);
Uint32 SDL_ThreadID(void);

/* mutexes / condition variables */
typedef struct SDL_mutex SDL_mutex;
typedef struct SDL_cond SDL_cond;
SDL_mutex* SDL_CreateMutex(void);
void SDL_DestroyMutex(SDL_mutex* m);
int SDL_mutexP(SDL_mutex* m); /* tolerates NULL: the reference never creates render_lock */
int SDL_mutexV(SDL_mutex* m);
#define SDL_LockMutex(m)   SDL_mutexP(m)
#define SDL_UnlockMutex(m) SDL_mutexV(m)
SDL_cond* SDL_CreateCond(void);
void SDL_DestroyCond(SDL_cond* c);
int SDL_CondWait(SDL_cond* c, SDL_mutex* m);
int SDL_CondSignal(SDL_cond* c);
int SDL_CondBroadcast(SDL_cond* c);

#ifdef __cplusplus
}
#endif

#endif
