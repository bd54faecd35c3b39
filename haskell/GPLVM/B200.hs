-- | FFI binding of libgplvm_b200.so (C ABI: include/gplvm_b200.h) for the reference's hot path.
--
-- UNVERIFIED: there is no GHC in the build image.  What IS checked mechanically (tests/test_haskell_shim.py): every
-- `foreign import ccall` below names a symbol exported by the library and matches its C prototype in arity and
-- argument / result types; every name in the export list is defined; the patches under haskell/patches apply to the
-- reference checkout.
--
-- Drop-in use: haskell/patches/*.patch replace the BODIES of emStepsFast / emStepsMissed (src/GPLVM/PPCA.hs:54-182,
-- src/GPLVM/TypeSafe/PPCA.hs:91-419), kernelFunction (src/GPLVM/GPExample.hs:93-100) and the cholSH / linearSolveS
-- calls of gpToPosteriorSample (src/GPLVM/GaussianProcess.hs:67-86) by calls into this module; every signature,
-- record and error text of the reference stays.  The library is deterministic for a fixed GPU count, which is what
-- makes `unsafePerformIO` under the reference's pure signatures sound.
{-# LANGUAGE ForeignFunctionInterface #-}
module GPLVM.B200
  ( ppcaDense
  , ppcaMissing
  , ppca
  , rbfKernel
  , gpCholLml
  , cholesky
  , triSolveLower
  , gpPosterior
  , pca
  , setDevices
  ) where

import Prelude

import Data.Array.Repa (Array, DIM1, DIM2, Z (..), (:.) (..), extent)
import Data.Array.Repa.Repr.ForeignPtr (F, fromForeignPtr, toForeignPtr)
import Foreign
import Foreign.C.String (peekCString)
import Foreign.C.Types
import System.IO.Unsafe (unsafePerformIO)

-- `safe`: the calls run for seconds under the -threaded RTS (GPLVMHaskell.cabal:84) and must not block other
-- Haskell threads or the garbage collector.
foreign import ccall safe "gplvm_ppca_dense"
  c_ppca_dense :: Ptr Double -> Int64 -> Int64 -> Int64 -> Ptr Double -> Double -> CInt -> Int64 -> Double
               -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Int64 -> IO CInt
foreign import ccall safe "gplvm_ppca_missing"
  c_ppca_missing :: Ptr Double -> Int64 -> Int64 -> Int64 -> Ptr Double -> Double -> CInt -> Int64 -> Double
                 -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Int64 -> IO CInt
foreign import ccall safe "gplvm_ppca"
  c_ppca :: Ptr Double -> Int64 -> Int64 -> Int64 -> Ptr Double -> Double -> CInt -> Int64 -> Double
         -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr CInt -> Ptr Int64 -> IO CInt
foreign import ccall safe "gplvm_rbf_kernel"
  c_rbf_kernel :: Ptr Double -> Int64 -> Ptr Double -> Int64 -> Int64 -> Ptr Double -> Double -> Ptr Double -> IO CInt
foreign import ccall safe "gplvm_gp_chol_lml"
  c_gp_chol_lml :: Ptr Double -> Int64 -> Int64 -> Ptr Double -> Ptr Double -> Double -> Double
                -> Ptr Double -> Ptr Double -> Ptr Double -> IO CInt
foreign import ccall safe "gplvm_cholesky"
  c_cholesky :: Ptr Double -> Int64 -> Ptr Double -> IO CInt
foreign import ccall safe "gplvm_tri_solve_lower"
  c_tri_solve_lower :: Ptr Double -> Int64 -> Ptr Double -> Int64 -> Ptr Double -> IO CInt
foreign import ccall safe "gplvm_gp_posterior"
  c_gp_posterior :: Ptr Double -> Int64 -> Ptr Double -> Ptr Double -> Int64 -> Int64 -> Ptr Double -> Double
                 -> Double -> Double -> Ptr Double -> Int64 -> Ptr Double -> Ptr Double -> Ptr Double -> IO CInt
foreign import ccall safe "gplvm_pca"
  c_pca :: Ptr Double -> Int64 -> Int64 -> Int64 -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Double
        -> Ptr Double -> Ptr Double -> Ptr CInt -> IO CInt
foreign import ccall safe "gplvm_set_devices"
  c_set_devices :: CInt -> IO CInt
foreign import ccall unsafe "gplvm_last_error"
  c_last_error :: IO (Ptr CChar)

-- | enum gplvm_status: 4 = GPLVM_ERR_NUMERIC (matrix not positive definite / singular -- where hmatrix would throw
-- or linearSolveS would return Nothing)
errNumeric :: CInt
errNumeric = 4

lastError :: IO String
lastError = c_last_error >>= peekCString

check :: CInt -> IO ()
check 0 = pure ()
check _ = lastError >>= \msg -> error ("libgplvm_b200: " ++ msg)   -- hmatrix throws at the same places

stopArgs :: Either Int Double -> (CInt, Int64, Double)   -- Either Int Double of src/GPLVM/PPCA.hs:36-37
stopArgs (Left k)    = (0, fromIntegral k, 0)
stopArgs (Right tol) = (1, 0, tol)

-- | Body of emStepsFast (src/GPLVM/PPCA.hs:54-94): (W, variance, expLogLikelihood).  y is D x N, w0 is D x M.
ppcaDense :: Array F DIM2 Double -> Array F DIM2 Double -> Double -> Either Int Double
          -> (Array F DIM2 Double, Double, Double)
ppcaDense y w0 s20 stop = unsafePerformIO $ do
  let Z :. d :. n = extent y
      Z :. _ :. m = extent w0
      (kind, k, tol) = stopArgs stop
  wOut <- mallocForeignPtrArray (d * m)
  withForeignPtr (toForeignPtr y) $ \py -> withForeignPtr (toForeignPtr w0) $ \pw0 ->
    withForeignPtr wOut $ \pw -> alloca $ \ps2 -> alloca $ \pll -> alloca $ \pst -> do
      check =<< c_ppca_dense py (fromIntegral d) (fromIntegral n) (fromIntegral m) pw0 s20 kind k tol pw ps2 pll pst
      s2 <- peek ps2
      ll <- peek pll
      pure (fromForeignPtr (Z :. d :. m) wOut, s2, ll)

-- | Body of emStepsMissed (src/GPLVM/PPCA.hs:96-182): (W, variance, expLogLikelihood, restored D x N).
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
      s2 <- peek ps2
      ll <- peek pll
      pure (fromForeignPtr (Z :. d :. m) wOut, s2, ll, fromForeignPtr (Z :. d :. n) rOut)

-- | The NaN dispatch of makePPCA (src/GPLVM/PPCA.hs:44-48) done inside the library:
-- (noMissedData, W, variance, expLogLikelihood, Maybe restored).
ppca :: Array F DIM2 Double -> Array F DIM2 Double -> Double -> Either Int Double
     -> (Bool, Array F DIM2 Double, Double, Double, Maybe (Array F DIM2 Double))
ppca y w0 s20 stop = unsafePerformIO $ do
  let Z :. d :. n = extent y
      Z :. _ :. m = extent w0
      (kind, k, tol) = stopArgs stop
  wOut <- mallocForeignPtrArray (d * m)
  rOut <- mallocForeignPtrArray (d * n)
  withForeignPtr (toForeignPtr y) $ \py -> withForeignPtr (toForeignPtr w0) $ \pw0 ->
    withForeignPtr wOut $ \pw -> withForeignPtr rOut $ \pr -> alloca $ \ps2 -> alloca $ \pll -> alloca $ \pnm ->
      alloca $ \pst -> do
        check =<< c_ppca py (fromIntegral d) (fromIntegral n) (fromIntegral m) pw0 s20 kind k tol pw ps2 pll pr pnm pst
        s2 <- peek ps2
        ll <- peek pll
        nm <- peek pnm
        let restored = if nm /= 0 then Nothing else Just (fromForeignPtr (Z :. d :. n) rOut)
        pure (nm /= 0, fromForeignPtr (Z :. d :. m) wOut, s2, ll, restored)

-- | sqDist + kernelFunction (src/GPLVM/GPExample.hs:79-100) generalised to Q-dimensional ARD inputs: x1 is n1 x Q,
-- x2 is n2 x Q, the result is n2 x n1 (the orientation of sqDist).  The reference is Q = 1, [1 / 0.1], variance 1.
rbfKernel :: Array F DIM2 Double -> Array F DIM2 Double -> [Double] -> Double -> Array F DIM2 Double
rbfKernel x1 x2 invL2 variance = unsafePerformIO $ do
  let Z :. n1 :. q = extent x1
      Z :. n2 :. _ = extent x2
  kOut <- mallocForeignPtrArray (n1 * n2)
  withForeignPtr (toForeignPtr x1) $ \p1 -> withForeignPtr (toForeignPtr x2) $ \p2 ->
    withArray invL2 $ \pl -> withForeignPtr kOut $ \pk ->
      check =<< c_rbf_kernel p1 (fromIntegral n1) p2 (fromIntegral n2) (fromIntegral q) pl variance pk
  pure (fromForeignPtr (Z :. n2 :. n1) kOut)

-- | cholSH (K + jitter I) of src/GPLVM/GaussianProcess.hs:67-68 fused with the kernel build, plus the GP log marginal
-- likelihood: (lower factor L, lml, log det).  Nothing where the matrix is not positive definite.
gpCholLml :: Array F DIM2 Double -> Array F DIM1 Double -> [Double] -> Double -> Double
          -> Maybe (Array F DIM2 Double, Double, Double)
gpCholLml x y invL2 variance jitter = unsafePerformIO $ do
  let Z :. n :. q = extent x
  lOut <- mallocForeignPtrArray (n * n)
  withForeignPtr (toForeignPtr x) $ \px -> withForeignPtr (toForeignPtr y) $ \py ->
    withArray invL2 $ \pl -> withForeignPtr lOut $ \pL -> alloca $ \plml -> alloca $ \pld -> do
      rc <- c_gp_chol_lml px (fromIntegral n) (fromIntegral q) py pl variance jitter pL plml pld
      if rc == errNumeric
        then pure Nothing
        else do
          check rc
          lml <- peek plml
          ld <- peek pld
          pure (Just (fromForeignPtr (Z :. n :. n) lOut, lml, ld))

-- | cholSH (src/GPLVM/GaussianProcess.hs:67,83) with the factor taken as LOWER (A = L L^T).  Nothing = not positive
-- definite (hmatrix throws there).
cholesky :: Array F DIM2 Double -> Maybe (Array F DIM2 Double)
cholesky a = unsafePerformIO $ do
  let Z :. n :. _ = extent a
  lOut <- mallocForeignPtrArray (n * n)
  rc <- withForeignPtr (toForeignPtr a) $ \pa -> withForeignPtr lOut $ \pL -> c_cholesky pa (fromIntegral n) pL
  if rc == errNumeric then pure Nothing else check rc >> pure (Just (fromForeignPtr (Z :. n :. n) lOut))

-- | linearSolveS cholK b (src/GPLVM/GaussianProcess.hs:74,77) for a lower-triangular factor: solves L X = B.
-- Nothing = singular factor (the reference's Maybe).
triSolveLower :: Array F DIM2 Double -> Array F DIM2 Double -> Maybe (Array F DIM2 Double)
triSolveLower l b = unsafePerformIO $ do
  let Z :. n :. _ = extent l
      Z :. _ :. r = extent b
  xOut <- mallocForeignPtrArray (n * r)
  rc <- withForeignPtr (toForeignPtr l) $ \pL -> withForeignPtr (toForeignPtr b) $ \pb -> withForeignPtr xOut $ \px ->
    c_tri_solve_lower pL (fromIntegral n) pb (fromIntegral r) px
  if rc == errNumeric then pure Nothing else check rc >> pure (Just (fromForeignPtr (Z :. n :. r) xOut))

-- | gpToPosteriorSample (src/GPLVM/GaussianProcess.hs:51-98) for the RBF/ARD kernel, whole pipeline on the device:
-- observe (n_obs x Q), train (n_train x Q), y_train, inverse squared length scales, variance, the two jitters
-- (5e-5 and 1e-6 in the reference, :68,:85) and the standard normals (n_obs x n_samples; the reference draws them
-- from mkStdGen (-4), :95-98).  Returns (mean, posterior factor (lower), samples); Nothing where the reference's
-- Maybe is Nothing / cholSH throws.
gpPosterior :: Array F DIM2 Double -> Array F DIM2 Double -> Array F DIM1 Double -> [Double] -> Double -> Double -> Double
            -> Array F DIM2 Double
            -> Maybe (Array F DIM1 Double, Array F DIM2 Double, Array F DIM2 Double)
gpPosterior observe train yTrain invL2 variance jitterTrain jitterPost normals = unsafePerformIO $ do
  let Z :. nObs :. q = extent observe
      Z :. nTrain :. _ = extent train
      Z :. _ :. nSamples = extent normals
  meanOut <- mallocForeignPtrArray nObs
  postOut <- mallocForeignPtrArray (nObs * nObs)
  sampOut <- mallocForeignPtrArray (nObs * nSamples)
  rc <- withForeignPtr (toForeignPtr observe) $ \po -> withForeignPtr (toForeignPtr train) $ \pt ->
    withForeignPtr (toForeignPtr yTrain) $ \py -> withArray invL2 $ \pl -> withForeignPtr (toForeignPtr normals) $ \pn ->
      withForeignPtr meanOut $ \pm -> withForeignPtr postOut $ \pp -> withForeignPtr sampOut $ \ps ->
        c_gp_posterior po (fromIntegral nObs) pt py (fromIntegral nTrain) (fromIntegral q) pl variance jitterTrain jitterPost
                       pn (fromIntegral nSamples) pm pp ps
  if rc == errNumeric
    then pure Nothing
    else do
      check rc
      pure (Just (fromForeignPtr (Z :. nObs) meanOut, fromForeignPtr (Z :. nObs :. nObs) postOut,
                  fromForeignPtr (Z :. nObs :. nSamples) sampOut))

-- | makePCA (src/GPLVM/PCA.hs:35-62) for an N x D input with the observations as ROWS, keeping d eigenvectors:
-- (mean (D), covariance (D x D), eigenvalues (D, descending), eigenvectors (D x d, columns), finalData (N x d),
-- restoredData (N x D)).  meanCovS / eigSH / mulP / pinvS of the reference all run inside the one call.
pca :: Array F DIM2 Double -> Int
    -> (Array F DIM1 Double, Array F DIM2 Double, Array F DIM1 Double, Array F DIM2 Double, Array F DIM2 Double, Array F DIM2 Double)
pca x dKeep = unsafePerformIO $ do
  let Z :. n :. dFeat = extent x
      d = min dKeep dFeat
  meanOut <- mallocForeignPtrArray dFeat
  covOut <- mallocForeignPtrArray (dFeat * dFeat)
  valsOut <- mallocForeignPtrArray dFeat
  vecsOut <- mallocForeignPtrArray (dFeat * d)
  finalOut <- mallocForeignPtrArray (n * d)
  restOut <- mallocForeignPtrArray (n * dFeat)
  withForeignPtr (toForeignPtr x) $ \px -> withForeignPtr meanOut $ \pm -> withForeignPtr covOut $ \pc ->
    withForeignPtr valsOut $ \pv -> withForeignPtr vecsOut $ \pe -> withForeignPtr finalOut $ \pf ->
      withForeignPtr restOut $ \pr -> alloca $ \psw ->
        check =<< c_pca px (fromIntegral n) (fromIntegral dFeat) (fromIntegral dKeep) pm pc pv pe pf pr psw
  pure ( fromForeignPtr (Z :. dFeat) meanOut, fromForeignPtr (Z :. dFeat :. dFeat) covOut, fromForeignPtr (Z :. dFeat) valsOut
       , fromForeignPtr (Z :. dFeat :. d) vecsOut, fromForeignPtr (Z :. n :. d) finalOut, fromForeignPtr (Z :. n :. dFeat) restOut )

-- | Number of GPUs of this process the library shards observations over (include/gplvm_b200.h: gplvm_set_devices).
setDevices :: Int -> IO ()
setDevices n = check =<< c_set_devices (fromIntegral n)
