--[[ mgconv_ffi.lua -- LuaJIT FFI binding of libmgconv.so (the C ABI in include/mgconv.h).

   The header is consumed verbatim: ffi.cdef() gets the text of include/mgconv.h with the
   preprocessor lines stripped, so this file can never drift from the ABI.  Usage:

       local mg = require 'mgconv_ffi'          -- package.path must contain <repo>/lua/?.lua
       local ctx = mg.ctx(cutorch.getDevice() - 1, cutorch.getStream-pointer, mg.C.MG_BF16)
       mg.check(ctx, mg.C.mg_pool_forward(ctx, gin, gout, 0, nil))

   NOTE: there is no LuaJIT / Torch7 in the build container nor on the GPU box (probed:
   `which luajit th lua` finds nothing), so this file is exercised only by inspection; the same
   header is bound and tested from Python ctypes (multigrid-neural-architectures_b200/mgconv/ffi.py),
   whose call sequence (engine.py / ops.py) is what lua/mgconv_nn.lua mirrors module by module.
]]
local ffi = require 'ffi'

local function script_dir()
   local src = debug.getinfo(1, 'S').source
   return (src:sub(1, 1) == '@' and src:sub(2) or src):match('(.*/)') or './'
end

local root = os.getenv('MGCONV_ROOT') or (script_dir() .. '../')
local function read(path)
   local f = assert(io.open(path, 'r'), 'mgconv: cannot open ' .. path)
   local s = f:read('*a'); f:close(); return s
end

-- strip #include / #ifdef / #define lines and the extern "C" braces; keep declarations + comments
local header = read(root .. 'include/mgconv.h')
local decl = {}
for line in header:gmatch('[^\n]*') do
   if not line:match('^%s*#') and not line:match('^extern "C"') and not line:match('^}%s*$') then
      decl[#decl + 1] = line
   end
end
-- the header keeps its array bounds (MG_MAX_SEG, MG_MAX_SRC) as enum constants, not macros, precisely
-- so that this text parses under LuaJIT's preprocessor-less ffi.cdef (and Python cffi, see
-- tests/test_host_cpu.py::test_header_parses_in_ffi_cdef_syntax)
ffi.cdef(table.concat(decl, '\n'))

local lib_path = os.getenv('MGCONV_LIB') or (root .. 'multigrid-neural-architectures_b200/mgconv/libmgconv.so')
local C = ffi.load(lib_path)   -- raises if missing: there is no CPU fallback for the hot path

local M = { C = C, ffi = ffi }

function M.check(ctx, status)
   if status ~= 0 then
      error(string.format('mgconv: status %d: %s', status, ffi.string(C.mg_last_error(ctx))), 2)
   end
end

--- one context per (GPU, Lua state): the reference runs one Lua state per GPU (multigpu.lua:94-98)
function M.ctx(device, stream, dtype)
   local out = ffi.new('mg_ctx*[1]')
   local rc = C.mg_ctx_create(device, stream, dtype or C.MG_BF16, out)
   if rc ~= 0 then error('mgconv: mg_ctx_create failed with status ' .. rc .. ' (no sm_100 GPU?)') end
   return ffi.gc(out[0], C.mg_ctx_destroy)
end

--- mg_grid view of a bf16/fp32 NHWC device buffer owned by Torch (raw pointer, never cached)
function M.grid(ptr, N, H, W, Cc, scale, shift, relu)
   local g = ffi.new('mg_grid')
   g.data = ptr; g.scale = scale; g.shift = shift; g.relu = relu or 0
   g.N, g.H, g.W, g.C, g.Cp = N, H, W, Cc, math.floor((Cc + 7) / 8) * 8
   return g
end

return M
