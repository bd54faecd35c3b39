-- UNVERIFIED (no GHC in the build image): FFI binding of libgplvm_b200.so, see INTEGRATION.md
{-# LANGUAGE ForeignFunctionInterface #-}
module GPLVM.B200 (ppcaDense, ppcaMissing, rbfKernel, gpCholLml) where

import Data.Array.Repa (Array, DIM2, Z (..), (:.) (..), computeS, extent)
import Data.Array.Repa.Repr.ForeignPtr (F, fromForeignPtr, toForeignPtr)
import Foreign
import Foreign.C.String (peekCString)
import Foreign.C.Types
import System.IO.Unsafe (unsafePerformIO)

-- include/gplvm_b200.h; `safe`: calls run for seconds under the -threaded RTS (GPLVMHaskell.cabal:84)
foreign import ccall safe "gplvm_ppca_dense"
  c_ppca_dense :: Ptr Double -> Int64 -> Int64 -> Int64 -> Ptr Double -> Double -> CInt -> Int64 -> Double
               -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Int64 -> IO CInt
foreign import ccall safe "gplvm_ppca_missing"
  c_ppca_missing :: Ptr Double -> Int64 -> Int64 -> Int64 -> Ptr Double -> Double -> CInt -> Int64 -> Double
                 -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Int64 -> IO CInt
foreign import ccall safe "gplvm_rbf_kernel"
  c_rbf_kernel :: Ptr Double -> Int64 -> Ptr Double -> Int64 -> Int64 -> Ptr Double -> Double -> Ptr Double -> IO CInt
foreign import ccall safe "gplvm_gp_chol_lml"
  c_gp_chol_lml :: Ptr Double -> Int64 -> Int64 -> Ptr Double -> Ptr Double -> Double -> Double
                -> Ptr Double -> Ptr Double -> Ptr Double -> IO CInt
foreign import ccall unsafe "gplvm_last_error" c_last_error :: IO (Ptr CChar)

check :: CInt -> IO ()
check 0 = pure ()
check _ = c_last_error >>= peekCString >>= \msg -> error ("libgplvm_b200: " <> msg)  -- hmatrix throws here too

stopArgs :: Either Int Double -> (CInt, Int64, Double)   -- Either Int Double of PPCA.hs:36-37
stopArgs (Left k)    = (0, fromIntegral k, 0)
stopArgs (Right tol) = (1, 0, tol)

-- | body of emStepsFast (PPCA.hs:54-94): (W, variance, expLogLikelihood)
ppcaDense :: Array F DIM2 Double -> Array F DIM2 Double -> Double -> Either Int Double
          -> (Array F DIM2 Double, Double, Double)
ppcaDense y w0 s20 stop = unsafePerformIO $ do      -- pure signature above => the library is deterministic
  let Z :. d :. n = extent y
      Z :. _ :. m = extent w0
      (kind, k, tol) = stopArgs stop
  wOut <- mallocForeignPtrArray (d * m)
  withForeignPtr (toForeignPtr y) $ \py -> withForeignPtr (toForeignPtr w0) $ \pw0 ->
    withForeignPtr wOut $ \pw -> alloca $ \ps2 -> alloca $ \pll -> alloca $ \pst -> do
      check =<< c_ppca_dense py (fromIntegral d) (fromIntegral n) (fromIntegral m) pw0 s20 kind k tol pw ps2 pll pst
      s2 <- peek ps2; ll <- peek pll
      pure (fromForeignPtr (Z :. d :. m) wOut, s2, ll)

-- | body of emStepsMissed (PPCA.hs:96-182): (W, variance, expLogLikelihood, restored)
ppcaMissing :: Array F DIM2 Double -> Array F DIM2 Double -> Double -> Either Int Double
            -> (Array F DIM2 Double, Double, Double, Array F DIM2 Double)
ppcaMissing y w0 s20 stop = unsafePerformIO $ do
  let Z :. d :. n = extent y
      Z :. _ :. m = extent w0
      (kind, k, tol) = stopArgs stop
  wOut <- mallocForeignPtrArray (d * m)
  rOut <- mallocForeignPtrArray (d * n)
  withForeignPtr (toForeignPtr y) $ \py -> withForeignPtr (toForeignPtr w0) $ \pw0 ->
    withForeignPtr wOut $ \pw -> withForeignPtr rOut $ \pr -> alloca $ \ps2 -> alloca $ \pll -> alloca $ \pst -> do
      check =<< c_ppca_missing py (fromIntegral d) (fromIntegral n) (fromIntegral m) pw0 s20 kind k tol pw ps2 pll pr pst
      s2 <- peek ps2; ll <- peek pll
      pure (fromForeignPtr (Z :. d :. m) wOut, s2, ll, fromForeignPtr (Z :. d :. n) rOut)
