#pragma once
namespace boost { namespace program_options { class parsed_options; } }
