// cwa/Module.h -- mirror of the reference's Module registry (CoupledWaterAnimation/Module.h:6-32,
// Module.cpp:5-91): objects register themselves on construction and the app fans Init / Compute /
// ... out to all of them.  Pure host logic, no device work.
#pragma once

#include <list>

#include "Common.h"

class Module {
public:
    Module() { sAllModules().push_back(this); }
    virtual ~Module() { sAllModules().remove(this); }
    virtual void Init() {}
    virtual void Reinit() {}
    virtual void Draw() {}
    virtual void DrawGui() {}
    virtual void Compute() {}
    virtual void Animate(float /*t*/ = -1.0f, float /*dt*/ = -1.0f) {}
    virtual void Keyboard(int, int, int, int) {}
    virtual void MouseCursor(cwa::vec2) {}
    virtual void MouseButton(int, int, int, cwa::vec2) {}

    static std::list<Module*>& sAllModules() { static std::list<Module*> all; return all; }
    static void sInitAll() { for (Module* m : sAllModules()) m->Init(); }
    static void sReinitAll() { for (Module* m : sAllModules()) m->Reinit(); }
    static void sDrawAll() { for (Module* m : sAllModules()) m->Draw(); }
    static void sDrawGuiAll() { for (Module* m : sAllModules()) m->DrawGui(); }
    static void sComputeAll() { for (Module* m : sAllModules()) m->Compute(); }
    static void sAnimateAll(float t, float dt) { for (Module* m : sAllModules()) m->Animate(t, dt); }
    static void sKeyboardAll(int k, int s, int a, int mo) { for (Module* m : sAllModules()) m->Keyboard(k, s, a, mo); }
    static void sMouseCursorAll(cwa::vec2 p) { for (Module* m : sAllModules()) m->MouseCursor(p); }
    static void sMouseButtonAll(int b, int a, int mo, cwa::vec2 p) { for (Module* m : sAllModules()) m->MouseButton(b, a, mo, p); }
};
