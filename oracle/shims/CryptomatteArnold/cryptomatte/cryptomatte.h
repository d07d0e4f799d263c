// Stand-in for CryptomatteArnold's header (absent third party): lentil.h:256-262 only reads one flag.
#pragma once
struct CryptomatteData { bool is_setup_completed = true; };
